"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck): both coarse paths on ragged
shapes, both thresholds paths, gather (channels-last and NCHW), fine match (fused and unfused), retrieval, host pipeline."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pope_b200 import _lib, driver, ops, synth
dev = torch.device("cuda:0")
for (n, hw0, hw1, C, dtype, impl, thr) in [(2, (12, 16), (10, 9), 256, torch.bfloat16, _lib.COARSE_TCGEN05, 0.2),
                                           (1, (20, 24), (18, 28), 128, torch.bfloat16, _lib.COARSE_TCGEN05, 0.1),
                                           (2, (12, 16), (10, 9), 64, torch.float32, _lib.COARSE_SIMT, 0.2),
                                           (2, (13, 17), (10, 9), 192, torch.float32, _lib.COARSE_TCGEN05, 0.2),   # split path
                                           (1, (37, 41), (33, 29), 256, torch.bfloat16, _lib.COARSE_TCGEN05, 0.3)]:
    L, S = hw0[0] * hw0[1], hw1[0] * hw1[1]
    f0, f1 = synth.coarse_features(5, n, L, S, C, sigma=0.8, dtype=dtype)
    ff0, _ = synth.fine_feature_maps(6, n, hw0[0] * 4, hw0[1] * 4, 128, dtype=dtype)
    ff1, _ = synth.fine_feature_maps(7, n, hw1[0] * 4, hw1[1] * 4, 128, dtype=dtype)
    for fused in (True, False):
        res = ops.match_pairs_device(f0.to(dev), f1.to(dev), ff0.to(dev), ff1.to(dev), (hw0[0] * 8, hw0[1] * 8), hw0, hw1,
                                     thr=thr, impl=impl, fused_fine=fused)
        print("impl", impl, "thr", thr, "fused", fused, "M", res.total(), "flags", res.flags())
    m = res.total()
    w0, w1 = ops.fine_gather(ff0.contiguous().to(dev), ff1.contiguous().to(dev), res["b_ids"][:m], res["i_ids"][:m],
                             res["j_ids"][:m], hw0[1], hw1[1], 4, 5)     # NCHW path
    order = ops.match_order_by_ref(res["counts"], n, S, res["j_ids"])
# fine-level transformer + FinePreprocess Linears on a ragged window count, match-list consumer, record packing
import pope_b200
torch.manual_seed(1)
mm = pope_b200.Matcher(pope_b200.make_default_cfg()).eval()
packed = torch.cat([ops.pack_fine_layer(l.state_dict(), dev) for l in mm.loftr_fine.layers])
for mw in (1, 37, 333):
    a = torch.randn(mw, 25, 128, device=dev).to(torch.bfloat16)
    b = torch.randn(mw, 25, 128, device=dev).to(torch.bfloat16)
    ops.fine_transformer(a, b, packed, mm.loftr_fine.layer_names)
    fc0 = torch.randn(2, 50, 256, device=dev).to(torch.bfloat16)
    fc1 = torch.randn(2, 40, 256, device=dev).to(torch.bfloat16)
    ids = [torch.randint(0, hi, (mw,), device=dev) for hi in (2, 50, 40)]
    ops.fine_merge_coarse(a, b, fc0, fc1, *ids, ops.pack_fine_pre(mm.fine_preprocess.state_dict(), dev))
    print("fine_tf windows", mw, "finite", bool(torch.isfinite(a.float()).all() and torch.isfinite(b.float()).all()))
print("match_scores", [t.tolist() for t in ops.match_scores(res["mconf"], res["counts"], n, group=2)])
print("pack_records", tuple(driver.pack_records(dict(res, mkpts0_f=res["mkpts0_c"], mkpts1_f=res["mkpts1_c"]), 3).shape))
# batched pose RANSAC on ragged match lists (one pair below five matches, one empty)
import numpy as np
from oracle.gen_golden_pose import scene
from pope_b200 import pose
rng = np.random.default_rng(0)
sc = [scene(rng, mm_, 0.3, 0.1, 600.0, 650.0) for mm_ in (300, 3, 0, 77, 1000)]
po = pose.estimate_pose_batch(torch.from_numpy(np.concatenate([s_[0] for s_ in sc])).to(dev),
                              torch.from_numpy(np.concatenate([s_[1] for s_ in sc])).to(dev),
                              torch.tensor([len(s_[0]) for s_ in sc], dtype=torch.int32, device=dev),
                              torch.from_numpy(np.stack([s_[2] for s_ in sc])), torch.from_numpy(np.stack([s_[3] for s_ in sc])),
                              0.5, 0.99999)
print("pose status", po["status"].tolist(), "inliers", po["n_inliers"].tolist(), "iters", po["iters"].tolist())
q, refs = synth.retrieval_tokens(3, 100, 384)
print("topk", ops.cosine_topk(q.to(dev), refs.to(dev), 3)[2].tolist())
f0, f1 = synth.coarse_features(8, 3, 192, 192, 256, sigma=0.8, dtype=torch.bfloat16)
ff0, ff1 = synth.fine_feature_maps(9, 3, 48, 64, 128, dtype=torch.bfloat16)
out = driver.match_pairs_host(f0, f1, ff0, ff1, (96, 128), (12, 16), (12, 16), chunk_pairs=2, device=0)
print("pipeline", int(out["counts"].sum()))
torch.cuda.synchronize()
print("sanitize_run done")
