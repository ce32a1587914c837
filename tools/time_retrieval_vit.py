"""Retrieval end to end for one query (BASELINE configs[2]: 1 query vs 256 reference crops at 224x224): the reference's
shape of the computation (R batch-1 ViT forwards, one F.cosine_similarity + .item() per crop, eval_linemod_json.py:74-101)
against pope_b200.retrieve_topk_images (batched forwards + device top-k).  Random-init ViT-S/14, fp32 and bf16."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
import pope_b200
from pope_b200.dino_vit import DinoViT

dev = torch.device("cuda:0")
R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
g = torch.Generator(device=dev).manual_seed(1)
for dtype in (torch.float32, torch.bfloat16):
    m = DinoViT().eval().to(dev, dtype)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() > 1:
                p.normal_(0, 0.05)
    ref_img = torch.randn(1, 3, 224, 224, device=dev, generator=g).to(dtype)
    crops = torch.randn(R, 3, 224, 224, device=dev, generator=g).to(dtype)

    def loop():                                      # the eval loop's structure
        with torch.no_grad():
            ref_fea = m(ref_img, is_training=True)["x_norm_clstoken"]
            sim, idx = np.array([0, 0, 0], np.float32), [-1, -1, -1]
            for r in range(R):
                fea = m(crops[r:r + 1], is_training=True)["x_norm_clstoken"]
                score = F.cosine_similarity(ref_fea, fea, dim=1, eps=1e-8)
                if (score.item() > sim).any():
                    k = int(np.argmin(sim)); sim[k] = score.item(); idx[k] = r
        return idx

    def batched():
        return pope_b200.retrieve_topk_images(m, ref_img, crops, k=3, batch=128)[2]

    for fn in (loop, batched):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            out = fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print(f"{str(dtype)[6:]:9s} {fn.__name__:8s}: {1e3 * dt:8.2f} ms per query (R = {R}), slots {out}")
