"""Randomised parity soak of the coarse + fine path against the oracle: random ragged shapes, channel counts, dtypes,
thresholds, border widths, temperatures, feature statistics and implementations.  Index sets must agree except at near-ties
(tests/parity_utils.py), confidences and fine coordinates within the dtype's tolerance.
    python tools/fuzz_parity.py [cases] [seed] [max_cells_per_side=40]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import pope_oracle as O
from pope_b200 import _lib, ops, synth
from tests.parity_utils import compare_match_lists, oracle_with_margins

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 40
dev = torch.device("cuda:0")
t0 = time.time()
tot = dict(cases=0, same=0, near=0, bad=0, conf_fail=0, fine_fail=0, flagged=0, empty=0, robust_path=0)
for case in range(cases):
    n = int(rng.integers(1, 4))
    h0, w0, h1, w1 = (int(rng.integers(3, hi + 1)) for _ in range(4))
    C = int(rng.choice([64, 128, 192, 256]))
    dtype = torch.bfloat16 if rng.random() < 0.6 else torch.float32
    thr = float(rng.choice([0.05, 0.1, 0.2, 0.3, 0.5, 0.9]))
    border = int(rng.integers(0, 3))
    temp = float(rng.choice([0.1, 0.1, 0.2, 0.05]))
    # a third of the cases with large norms (token norm up to 64: similarities in the hundreds of log2 units at T = 0.05)
    sigma = float(rng.uniform(0.5, 1.3)) if rng.random() < 0.67 else float(rng.uniform(1.3, 4.0))
    noise = float(rng.choice([0.0, 0.1, 0.3, 0.6]))
    planted = float(rng.uniform(0.0, 1.0))
    impl = _lib.COARSE_SIMT if rng.random() < 0.25 else _lib.COARSE_AUTO
    L, S = h0 * w0, h1 * w1
    f0, f1 = synth.coarse_features(1000 + case, n, L, S, C, sigma=sigma, planted=planted, noise=noise, dtype=dtype)
    if rng.random() < 0.2:                                      # norms that grow / shrink along the sweep: the lazy shift must rise
        g1 = torch.logspace(0, float(rng.uniform(0.0, 0.5)), S)
        g0 = torch.logspace(float(rng.uniform(0.0, 0.3)), 0, L)
        f1 = (f1.float() * (g1 if rng.random() < 0.5 else g1.flip(0))[None, :, None]).to(dtype)
        f0 = (f0.float() * g0[None, :, None]).to(dtype)
    if rng.random() < 0.15:                                     # duplicated rows: exact ties
        f1[:, : S // 2] = f1[:, S - S // 2:]
    ff0, _ = synth.fine_feature_maps(2000 + case, n, h0 * 4, w0 * 4, 128, dtype=dtype)
    _, ff1 = synth.fine_feature_maps(3000 + case, n, h1 * 4, w1 * 4, 128, dtype=dtype)
    desc = f"case {case}: n={n} {h0}x{w0} vs {h1}x{w1} C={C} {str(dtype)[6:]} thr={thr} border={border} T={temp} sigma={sigma:.2f} impl={impl}"
    try:
        res = ops.match_pairs_device(f0.to(dev), f1.to(dev), ff0.to(dev), ff1.to(dev), (h0 * 8, w0 * 8), (h0, w0), (h1, w1),
                                     thr=thr, border_rm=border, temperature=temp, impl=impl)
        torch.cuda.synchronize()
    except Exception as e:                                      # an unsupported combination must be a clean error
        print(desc, "-> raised", type(e).__name__, str(e)[:80])
        continue
    flags = res.flags()
    if flags & ~_lib.FLAG_ROBUST_PATH:
        tot["flagged"] += 1
        print(desc, "-> flags", flags)
        continue
    tot["robust_path"] += int(bool(flags & _lib.FLAG_ROBUST_PATH))
    m = res.total()
    got = {k: res[k][:m].cpu() for k in ("b_ids", "i_ids", "j_ids", "mconf", "mkpts0_f", "mkpts1_f")}
    want, mg = oracle_with_margins(f0.float(), f1.float(), (h0 * 8, w0 * 8), (h0, w0), (h1, w1), thr=thr, border_rm=border,
                                   temperature=temp)
    same, near, bad = compare_match_lists(got, want, mg, thr=thr)
    tot["cases"] += 1; tot["same"] += same; tot["near"] += len(near); tot["bad"] += len(bad); tot["empty"] += int(m == 0)
    if bad:
        print(desc, "-> BAD", bad[:4])
    if not near and not bad and m:
        tol = 1e-2 if dtype == torch.bfloat16 else 2e-4
        if not torch.allclose(got["mconf"], want["mconf"], rtol=tol, atol=1e-7):
            tot["conf_fail"] += 1
            print(desc, "-> conf max rel err", float(((got["mconf"] - want["mconf"]).abs() / want["mconf"]).max()))
        full = O.match_pairs(f0.float(), f1.float(), ff0.float(), ff1.float(), (h0 * 8, w0 * 8), (h0, w0), (h1, w1), thr=thr,
                             border_rm=border, temperature=temp)
        # the batched oracle may break an exact tie (duplicated rows) differently from the per-pair oracle used above
        if not torch.equal(full["j_ids"], got["j_ids"]):
            tot["near"] += 1
        elif not (torch.allclose(got["mkpts1_f"], full["mkpts1_f"], rtol=tol, atol=2e-3 if dtype == torch.float32 else 5e-2)
                  and torch.equal(got["mkpts0_f"], full["mkpts0_f"])):
            tot["fine_fail"] += 1
            print(desc, "-> fine max abs err", float((got["mkpts1_f"] - full["mkpts1_f"]).abs().max()))
print(f"fuzz: {tot} in {time.time() - t0:.0f} s")
sys.exit(1 if tot["bad"] or tot["conf_fail"] or tot["fine_fail"] or tot["flagged"] else 0)
