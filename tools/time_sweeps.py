"""Times the coarse stage (64 pairs, 480x640, bf16) with CUDA events; used with POPE_TC_DEBUG experiments."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pope_b200 import _lib, ops, synth
n = 64
dev = torch.device("cuda:0")
f0, f1 = synth.coarse_features(1234, n, 4800, 4800, 256, dtype=torch.bfloat16)
d0, d1 = f0.to(dev), f1.to(dev)
ws = torch.empty(_lib.lib().pope_coarse_workspace_bytes(n, 4800, 4800), dtype=torch.uint8, device=dev)
for _ in range(3):
    r = ops.coarse_match(d0, d1, (60, 80), (60, 80), 8.0, workspace=ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    r = ops.coarse_match(d0, d1, (60, 80), (60, 80), 8.0, workspace=ws)
e1.record(); torch.cuda.synchronize()
print("POPE_TC_DEBUG=%s coarse %.3f ms/step  M=%d" % (os.environ.get("POPE_TC_DEBUG", "0"), e0.elapsed_time(e1) / 10, r.total()))
