"""One hot-path step (64 pairs at 480x640, bf16, device-resident) launched eagerly vs replayed as a CUDA graph, single
stream, CUDA events.   python tools/time_graph_step.py [pairs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pope_b200 import ops, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
hw = (60, 80)
f0, f1 = synth.coarse_features(1, n, 4800, 4800, 256, dtype=torch.bfloat16)
ff0, ff1 = synth.fine_feature_maps(2, n, 240, 320, 128, dtype=torch.bfloat16)
f0, f1, ff0, ff1 = f0.to(dev), f1.to(dev), ff0.to(dev), ff1.to(dev)
res = ops.match_pairs_device(f0, f1, ff0, ff1, (480, 640), hw, hw)
ws = res["workspace"]
for _ in range(3):
    res = ops.match_pairs_device(f0, f1, ff0, ff1, (480, 640), hw, hw, workspace=ws)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    gres = ops.match_pairs_device(f0, f1, ff0, ff1, (480, 640), hw, hw, workspace=ws)
graph.replay()
torch.cuda.synchronize()
m = res.total()
same = gres.total() == m and all(torch.equal(gres[k][:m], res[k][:m]) for k in ("b_ids", "i_ids", "j_ids", "mconf", "mkpts1_f"))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
out = {}
for name, fn in (("eager", lambda: ops.match_pairs_device(f0, f1, ff0, ff1, (480, 640), hw, hw, workspace=ws)), ("graph", graph.replay)):
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    out[name] = e0.elapsed_time(e1) / reps
print(f"{n} pairs, {m} matches: eager {out['eager']:.4f} ms per step ({n / out['eager'] * 1e3:,.0f} pairs/s), CUDA graph replay "
      f"{out['graph']:.4f} ms ({n / out['graph'] * 1e3:,.0f} pairs/s); same result: {same}")
