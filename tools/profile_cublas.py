"""Calibration for the tensor-pipe metric: what does sm__pipe_tensor_cycles_active read for a cuBLAS bf16 GEMM
(8192^3, the MEASURED_PEAKS workload) on this machine?  Run under ncu -k regex:gemm|nvjet|cutlass."""
import torch
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    c = a @ b
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    c = a @ b
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("cublas bf16 8192^3: %.3f ms  %.1f TFLOP/s" % (ms, 2 * 8192 ** 3 / ms / 1e9))
