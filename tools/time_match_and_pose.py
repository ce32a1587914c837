"""The whole device-resident chain on geometry-consistent synthetic pairs (synth.posed_pair_features): Matcher hot path
(coarse match -> fused fine match) -> batched pose RANSAC, timed per stage with CUDA events, pose checked against the
planted motion.   python tools/time_match_and_pose.py [pairs] [thresh_px] [conf]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pope_b200 import ops, pose, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
conf = float(sys.argv[3]) if len(sys.argv) > 3 else 0.99999
dev = torch.device("cuda:0")
hw = (60, 80)
d = synth.posed_pair_features(7, n, hw_c=hw, dtype=torch.bfloat16)
fc0, fc1, ff0, ff1 = (d[k].to(dev) for k in ("feat_c0", "feat_c1", "feat_f0", "feat_f1"))
K = d["K"].expand(n, 3, 3).to(dev).contiguous()
ws_c = ws_p = None


def step():
    global ws_c, ws_p
    res = ops.match_pairs_device(fc0, fc1, ff0, ff1, (480, 640), hw, hw, workspace=ws_c)
    ws_c = res.get("workspace", ws_c)
    out = pose.estimate_pose_batch(res["mkpts0_f"], res["mkpts1_f"], res["counts"], K, K, thresh, conf, workspace=ws_p)
    ws_p = out["workspace"]
    return res, out


for _ in range(3):
    res, out = step()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
reps, tm, tp = 10, [], []
for _ in range(reps):
    ev[0].record()
    res = ops.match_pairs_device(fc0, fc1, ff0, ff1, (480, 640), hw, hw, workspace=ws_c)
    ev[1].record()
    out = pose.estimate_pose_batch(res["mkpts0_f"], res["mkpts1_f"], res["counts"], K, K, thresh, conf, workspace=ws_p)
    ev[2].record()
    torch.cuda.synchronize()
    tm.append(ev[0].elapsed_time(ev[1])); tp.append(ev[1].elapsed_time(ev[2]))
R = out["R"].cpu().numpy()
errs = [float(np.degrees(np.arccos(np.clip((np.trace(R[p].T @ d["R"][p].numpy()) - 1) / 2, -1, 1)))) for p in range(n)]
m = res.total()
print(f"{n} pairs at 480x640 (bf16): {m} matches; match {np.median(tm):.3f} ms + pose {np.median(tp):.3f} ms = "
      f"{n / (np.median(tm) + np.median(tp)) * 1e3:,.0f} pairs/s for the chain; RANSAC thresh {thresh} px conf {conf}: iters median "
      f"{int(out['iters'].median())}, inlier frac {float(out['n_inliers'].sum()) / m:.3f}, status ok {int(out['status'].sum())}/{n}, "
      f"rotation error vs planted: median {np.median(errs):.3f} max {max(errs):.3f} deg")
