"""Randomised soak of pope_estimate_pose_batch against oracle/pose_oracle.py (bit for bit): random list lengths, outlier
ratios, noise, intrinsics, thresholds, confidences, iteration bounds and seeds, batched with ragged counts.
    python tools/fuzz_pose.py [batches] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import pose_oracle as O
from oracle.gen_golden_pose import scene
from pope_b200 import pose

batches = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = torch.device("cuda:0")
t0 = time.time()
pairs = mismatched = solved = 0
for bi in range(batches):
    n = int(rng.integers(1, 9))
    sc = []
    for _ in range(n):
        m = int(rng.choice([0, 3, 5, 6, 8, 20, 60, 150, 400, 900]))
        s = scene(rng, m, float(rng.choice([0.0, 0.2, 0.5, 0.8])), float(rng.choice([0.0, 0.05, 0.3, 1.0])),
                  float(rng.uniform(300, 900)), float(rng.uniform(300, 900)))
        if rng.random() < 0.1 and m:
            s = (np.repeat(s[0][:1], m, 0),) + s[1:]                # degenerate: every match at the same image-0 point
        sc.append(s)
    thresh = float(rng.choice([0.25, 0.5, 1.0, 3.0]))
    conf = float(rng.choice([0.9, 0.99, 0.99999, 1.0]))
    max_iters = int(rng.choice([1, 7, 64, 100, 300, 1000, 1024]))
    seed = int(rng.integers(0, 2 ** 40))
    mk0, mk1 = np.concatenate([s[0] for s in sc]), np.concatenate([s[1] for s in sc])
    counts = np.array([len(s[0]) for s in sc], dtype=np.int32)
    K0, K1 = np.stack([s[2] for s in sc]), np.stack([s[3] for s in sc])
    got = pose.estimate_pose_batch(torch.from_numpy(mk0).to(dev), torch.from_numpy(mk1).to(dev), torch.from_numpy(counts).to(dev),
                                   torch.from_numpy(K0), torch.from_numpy(K1), thresh, conf, max_iters, seed)
    want = O.estimate_pose_batch(mk0, mk1, counts, K0, K1, thresh, conf, max_iters, seed)
    pairs += n
    solved += int(want["status"].sum())
    for k in ("status", "iters", "n_inliers", "inliers", "E", "R", "t"):
        if not np.array_equal(got[k].cpu().numpy(), want[k], equal_nan=True):
            mismatched += 1
            print(f"batch {bi}: counts={counts.tolist()} thresh={thresh} conf={conf} max_iters={max_iters} seed={seed}: {k} differs")
            break
print(f"fuzz_pose: {batches} batches, {pairs} pairs ({solved} with a pose), {mismatched} batches differ from the oracle, {time.time() - t0:.0f} s")
sys.exit(1 if mismatched else 0)
