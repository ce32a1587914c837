"""Developer probe: does the fused fine kernel run beside the coarse sweep when they are on different streams?
Times coarse alone, fine alone, and both launched together on two streams."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pope_b200 import _lib, ops, synth
n = 64
dev = torch.device("cuda:0")
f0, f1 = synth.coarse_features(1234, n, 4800, 4800, 256, dtype=torch.bfloat16)
g = torch.Generator(device=dev).manual_seed(4321)
ff0 = torch.randn(n, 240, 320, 128, device=dev, generator=g).to(torch.bfloat16).permute(0, 3, 1, 2)
ff1 = torch.randn(n, 240, 320, 128, device=dev, generator=g).to(torch.bfloat16).permute(0, 3, 1, 2)
d0, d1 = f0.to(dev), f1.to(dev)
ws = [torch.empty(_lib.lib().pope_coarse_workspace_bytes(n, 4800, 4800), dtype=torch.uint8, device=dev) for _ in range(2)]
res = ops.coarse_match(d0, d1, (60, 80), (60, 80), 8.0, workspace=ws[0])
m_dev = res["counts"][n:n + 1]
torch.cuda.synchronize()
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()

def coarse():
    return ops.coarse_match(d0, d1, (60, 80), (60, 80), 8.0, workspace=ws[1])

def fine():
    return ops.fine_match_maps(ff0, ff1, res["b_ids"], res["i_ids"], res["j_ids"], res["mkpts1_c"], 80, 80, 4, 4.0, 5, m_dev)

def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def both():
    cur = torch.cuda.current_stream()
    sa.wait_stream(cur); sb.wait_stream(cur)
    with torch.cuda.stream(sb): coarse()
    with torch.cuda.stream(sa):
        torch.cuda._sleep(400000)        # ~0.2 ms: the sweep is resident on every SM when the fine kernel is launched
        fine()
    cur.wait_stream(sa); cur.wait_stream(sb)

print(f"coarse alone {timed(coarse):.3f} ms, fine alone {timed(fine):.3f} ms, both on two streams {timed(both):.3f} ms")
