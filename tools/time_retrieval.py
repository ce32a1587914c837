"""Times the retrieval kernels (BASELINE configs[2]: 1 query vs 256 reference crops): CLS-token variant (D = 384) and the
patch-token stress shape (256 tokens x 384 per crop, flattened to D = 98 304), bf16 and fp32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pope_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(5)
for name, D in (("CLS token", 384), ("patch tokens", 256 * 384)):
    for dt in (torch.float32, torch.bfloat16):
        q = torch.randn(1, D, device=dev, generator=g).to(dt)
        refs = torch.randn(256, D, device=dev, generator=g).to(dt)
        for _ in range(3):
            ops.cosine_topk(q, refs, 3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            s, ss, si = ops.cosine_topk(q, refs, 3)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        byts = (256 + 1) * D * q.element_size()
        print(f"{name:12s} {str(dt)[6:]:9s} D={D:6d}: {us:7.1f} us per query ({byts / us / 1e3:7.1f} GB/s), top-3 {si.tolist()}")
