"""Two steps of the bench workload (64 pairs at 480x640, bf16, device-resident) for ncu: step 1 warms up, step 2 is
the one to profile (`-s 7 -c 7` with -k regex:"sweep_tc|count_kernel|emit_kernel|gather_cl|fine_match")."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pope_b200 import _lib, ops, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
f0, f1 = synth.coarse_features(1234, n, 4800, 4800, 256, dtype=torch.bfloat16)
g = torch.Generator(device=dev).manual_seed(4321)
ff0 = torch.randn(n, 240, 320, 128, device=dev, generator=g).to(torch.bfloat16).permute(0, 3, 1, 2)
ff1 = torch.randn(n, 240, 320, 128, device=dev, generator=g).to(torch.bfloat16).permute(0, 3, 1, 2)
d0, d1 = f0.to(dev), f1.to(dev)
for _ in range(steps):
    res = ops.match_pairs_device(d0, d1, ff0, ff1, (480, 640), (60, 80), (60, 80))
torch.cuda.synchronize()
print("matches", res.total(), "flags", res.flags())
