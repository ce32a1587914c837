"""Steps 3-5 of Matcher.forward with fine_cuda_bf16=True on the bench workload (64 pairs, bf16), 3 calls: for an
`ncu --metrics gpu__time_duration.sum` launch list and for stage timings with CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pope_b200
from pope_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = pope_b200.Matcher(pope_b200.make_default_cfg(), fine_cuda_bf16=True).eval().to(dev)
f0, f1 = synth.coarse_features(99, n, 4800, 4800, 256, dtype=torch.bfloat16)
g = torch.Generator(device=dev).manual_seed(98)
ff0 = torch.randn(n, 240, 320, 128, device=dev, generator=g).to(torch.bfloat16).permute(0, 3, 1, 2)
ff1 = torch.randn(n, 240, 320, 128, device=dev, generator=g).to(torch.bfloat16).permute(0, 3, 1, 2)
f0, f1 = f0.to(dev), f1.to(dev)
shapes = {"hw0_i": torch.Size([480, 640]), "hw1_i": torch.Size([480, 640]), "hw0_c": torch.Size([60, 80]),
          "hw1_c": torch.Size([60, 80]), "hw0_f": torch.Size([240, 320]), "hw1_f": torch.Size([240, 320]), "bs": n}
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
for it in range(3):
    data = dict(shapes)
    with torch.no_grad():
        ev[0].record(); m.coarse_matching(f0, f1, data)
        ev[1].record(); w0, w1 = m.fine_preprocess(ff0, ff1, f0, f1, data)
        ev[2].record(); w0, w1 = m.loftr_fine(w0, w1)
        ev[3].record(); m.fine_matching(w0, w1, data)
        ev[4].record()
    torch.cuda.synchronize()
names = ["coarse_matching", "fine_preprocess (gather + Linears)", "loftr_fine", "fine_matching"]
print("matches", data["mconf"].numel(), " ".join(f"{nm}={ev[i].elapsed_time(ev[i + 1]):.2f}ms" for i, nm in enumerate(names)),
      f"total={ev[0].elapsed_time(ev[4]):.2f}ms")
