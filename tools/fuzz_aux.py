"""Randomised soak of the small kernels around the path: retrieval cosine + running top-k, and the match-list consumer
(pope_match_scores).  Scores are compared with the oracle within float tolerance; the slot logic is compared exactly by
running the oracle's slot loop on the scores the device produced.     python tools/fuzz_aux.py [cases] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import pope_oracle as O
from pope_b200 import ops

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = torch.device("cuda:0")
bad = 0
t0 = time.time()
for case in range(cases):
    R = int(rng.choice([1, 2, 3, 7, 64, 255, 256, 257, 1000, 3000]))
    D = int(rng.choice([8, 100, 384, 768, 1024, 1536]))
    k = int(rng.integers(1, 17))
    dtype = torch.float32 if rng.random() < 0.5 else torch.bfloat16
    g = torch.Generator().manual_seed(case)
    q = torch.randn(1, D, generator=g)
    refs = torch.randn(R, D, generator=g)
    hot = torch.randperm(R, generator=g)[: max(1, R // 8)]
    refs[hot] = q * float(rng.uniform(0.2, 3.0)) + float(rng.uniform(0.0, 2.0)) * torch.randn(hot.numel(), D, generator=g)
    if rng.random() < 0.2:
        refs[hot[: max(1, hot.numel() // 2)]] = refs[hot[0]].clone()          # exact score ties
    if rng.random() < 0.1:
        refs[0] = 0                                                    # zero vector: eps path of cosine_similarity
    q, refs = q.to(dtype), refs.to(dtype)
    scores, slot_s, slot_i = ops.cosine_topk(q.to(dev), refs.to(dev), k)
    want = O.cosine_scores(q, refs)
    sc = scores.cpu()
    ws, wi = O.running_topk(sc.tolist(), k)
    ok = torch.allclose(sc, want, rtol=2e-5, atol=2e-6) and slot_i.cpu().tolist() == wi and \
        torch.equal(slot_s.cpu(), torch.tensor(ws, dtype=torch.float32))
    if not ok:
        bad += 1
        print(f"retrieval case {case}: R={R} D={D} k={k} {dtype}: max score err {float((sc - want).abs().max()):.2e}, slots {slot_i.cpu().tolist()} vs {wi}")
    # match-list consumer
    n = int(rng.integers(1, 40))
    group = int(rng.integers(1, 6))
    thr = float(rng.choice([0.2, 0.5, 0.9, 0.99]))
    counts = rng.integers(0, 60, n).astype(np.int32)
    if rng.random() < 0.3:
        counts[rng.integers(0, n)] = 0
    conf = rng.random(int(counts.sum())).astype(np.float32)
    if rng.random() < 0.3:
        conf[:] = np.round(conf, 1)                                    # values exactly on thresholds, equal group scores
    per_pair = np.split(conf, np.cumsum(counts)[:-1])
    ws_, wb = O.match_scores(per_pair, group, thr)
    s_, b_ = ops.match_scores(torch.from_numpy(conf).to(dev), torch.from_numpy(counts).to(dev), n, group, thr)
    if s_.cpu().tolist() != ws_ or b_.cpu().tolist() != wb:
        bad += 1
        print(f"match_scores case {case}: n={n} group={group} thr={thr}: {s_.cpu().tolist()} vs {ws_}; {b_.cpu().tolist()} vs {wb}")
print(f"fuzz_aux: {cases} cases, {bad} failures, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
