"""Randomised soak of the host pipeline (pope_pipeline_run) with PAGE-LOCKED buffers against the device path, bit for bit:
random ragged shapes (different grids for the two images), chunk sizes, dtypes, thresholds, border widths 0-2 (border 0 puts
windows across the map's edge) and all three ways image 1's fine map can reach the device (POPE_PIPELINE_F1 = union /
windows / bulk).  For the union form the bytes the library reports are checked against the union of the matched cells'
windows computed here.      python tools/fuzz_pipeline.py [cases] [seed]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pope_b200 import driver, ops, synth

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = torch.device("cuda:0")
t0 = time.time()
bad = 0
rows = 0
for case in range(cases):
    n = int(rng.integers(1, 8))
    h0, w0, h1, w1 = (int(rng.integers(3, 30)) for _ in range(4))
    dtype = torch.bfloat16 if rng.random() < 0.6 else torch.float32
    esz = 2 if dtype == torch.bfloat16 else 4
    thr = float(rng.choice([0.1, 0.2, 0.3]))
    border = int(rng.integers(0, 3))
    chunk = int(rng.integers(1, n + 2))
    planted = float(rng.choice([0.2, 0.7, 1.0]))
    fc0, fc1 = synth.coarse_features(900 + case, n, h0 * w0, h1 * w1, 256, sigma=float(rng.uniform(0.7, 1.1)), planted=planted,
                                     dtype=dtype)
    ff0, _ = synth.fine_feature_maps(1000 + case, n, h0 * 4, w0 * 4, 128, dtype=dtype)
    _, ff1 = synth.fine_feature_maps(1100 + case, n, h1 * 4, w1 * 4, 128, dtype=dtype)
    res = ops.match_pairs_device(fc0.to(dev), fc1.to(dev), ff0.to(dev), ff1.to(dev), (h0 * 8, w0 * 8), (h0, w0), (h1, w1), thr=thr,
                                 border_rm=border)
    mt = res.total()
    rows += mt
    need = np.zeros((n, h1 * 4, w1 * 4), dtype=bool)
    for b, j in zip(res["b_ids"][:mt].tolist(), res["j_ids"][:mt].tolist()):
        cy, cx = divmod(j, w1)
        need[b, max(0, 4 * cy - 2):4 * cy + 3, max(0, 4 * cx - 2):4 * cx + 3] = True
    p0, p1 = fc0.pin_memory(), fc1.pin_memory()
    q0 = ff0.permute(0, 2, 3, 1).contiguous().pin_memory()
    q1 = ff1.permute(0, 2, 3, 1).contiguous().pin_memory()
    pl = driver.Pipeline(dtype, min(chunk, n), (h0 * 8, w0 * 8), (h0, w0), (h1, w1), thr=thr, border_rm=border, device=0)
    for mode in ("union", "windows", "bulk"):
        os.environ["POPE_PIPELINE_F1"] = mode
        out = pl.run(p0, p1, q0, q1)
        cat = driver.flatten_slots(out)
        ok = pl.last_f1_mode == mode and int(out["counts"].sum()) == mt and all(
            torch.equal(cat[k], res[k][:mt].cpu()) for k in ("b_ids", "i_ids", "j_ids", "mconf", "mkpts0_f", "mkpts1_f"))
        if mode == "union":
            want = (fc0.numel() + fc1.numel()) * esz + mt * 128 * esz + int(need.sum()) * 128 * esz
            ok = ok and pl.last_h2d_bytes == want
        if not ok:
            bad += 1
            print(f"case {case}: n={n} {h0}x{w0} vs {h1}x{w1} {str(dtype)[6:]} thr={thr} border={border} chunk={chunk} mode={mode}"
                  f" -> differs (matches {int(out['counts'].sum())} vs {mt}, bytes {pl.last_h2d_bytes})")
    del os.environ["POPE_PIPELINE_F1"]
    pl.close()
print(f"fuzz_pipeline: {cases} cases x 3 forms, {rows} match rows, {bad} failures, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
