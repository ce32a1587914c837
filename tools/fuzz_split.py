"""Randomised check of the head / tail launch form of the single sweep (csrc/coarse_tc.cu::coarse_tc_run, split_head): for
random batch sizes and ragged shapes the results must equal the one-launch form (POPE_TC_DEBUG=1024) bit for bit, in bf16 and
(three-way split path) fp32, also with hard-set pairs that the gated robust launch redoes.
    python tools/fuzz_split.py [cases] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pope_b200 import _lib, ops, synth

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = torch.device("cuda:0")
t0 = time.time()
tot = dict(cases=0, split=0, rows=0, flagged_cases=0, differ=0)
keys = ("b_ids", "i_ids", "j_ids", "mconf", "mkpts0_c", "mkpts1_c")
for case in range(cases):
    h0, w0, h1, w1 = (int(rng.integers(12, 57)) for _ in range(4))
    L, S = h0 * w0, h1 * w1
    upp = (L + 255) // 256
    n = int(rng.integers(2, max(3, min(96, 600 // upp))))
    dtype = torch.bfloat16 if rng.random() < 0.75 else torch.float32
    sigma = float(rng.uniform(0.7, 1.3)) if rng.random() < 0.7 else float(rng.uniform(1.3, 3.2))
    thr = float(rng.choice([0.2, 0.2, 0.3, 0.5]))
    f0, f1 = synth.coarse_features(5000 + case, n, L, S, 256, sigma=sigma, dtype=dtype)
    hard = rng.random() < 0.3
    if hard:                                  # one or two hard-set pairs anywhere in the batch
        k = int(rng.integers(1, 3))
        fh0, fh1 = synth.hard_coarse_features(6000 + case, k, L, S, 256, sigma=0.9, dtype=dtype)
        for q, where in enumerate(rng.choice(n, size=k, replace=False)):
            f0[int(where)], f1[int(where)] = fh0[q], fh1[q]
    d0, d1 = f0.to(dev), f1.to(dev)
    outs = []
    for knob in ("0", "1024"):
        os.environ["POPE_TC_DEBUG"] = knob
        res = ops.coarse_match(d0, d1, (h0, w0), (h1, w1), 8.0, thr=thr)
        torch.cuda.synchronize()
        m = res.total()
        outs.append((m, res.flags(), {k: res[k][:m].clone() for k in keys}))
    del os.environ["POPE_TC_DEBUG"]
    tot["cases"] += 1
    tot["split"] += int(_lib.single_sweep_is_split(n, L))
    tot["rows"] += outs[0][0]
    tot["flagged_cases"] += int(bool(outs[0][1] & _lib.FLAG_ROBUST_PATH))
    same = outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1] and all(torch.equal(outs[0][2][k], outs[1][2][k]) for k in keys)
    if not same or (outs[0][1] & ~_lib.FLAG_ROBUST_PATH):
        tot["differ"] += 1
        print(f"case {case}: n={n} {h0}x{w0} vs {h1}x{w1} {str(dtype)[6:]} sigma={sigma:.2f} thr={thr} hard={hard} -> "
              f"M {outs[0][0]} / {outs[1][0]}, flags {outs[0][1]} / {outs[1][1]}")
print(f"fuzz_split: {tot} in {time.time() - t0:.0f} s")
sys.exit(1 if tot["differ"] else 0)
