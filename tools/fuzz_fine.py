"""Randomised soak of (a) the bf16 fine-level transformer + FinePreprocess Linears against the fp32 oracle (rms error
relative to the output scale, the bound of tests/test_fine_tf.py) over window counts around every tile boundary and random
layer schedules, and (b) the host pipeline (pope_match_pairs_host, chunked streams) against the device path (exact) over
random ragged shapes, chunk sizes, dtypes and thresholds.      python tools/fuzz_fine.py [cases] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import pope_oracle as O
from pope_b200 import driver, ops, synth

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = torch.device("cuda:0")
t0 = time.time()
bad = 0


def rel(got, want):
    rms = want.pow(2).mean().sqrt()
    d = (got.float() - want).abs()
    return float(d.pow(2).mean().sqrt() / rms), float(d.max() / rms)


for case in range(cases):
    # ---- (a) fine transformer + merge
    m = int(rng.choice([1, 2, 5, 6, 10, 11, 20, 21, 41, 82, 127, 128, 129, 163, 164, 165, 500, 1515, 3031]))
    names = [str(rng.choice(["self", "cross"])) for _ in range(int(rng.integers(1, 5)))]
    gen = torch.Generator().manual_seed(case)
    layers = []
    for _ in names:
        sd = {}
        for key, shp in (("q_proj.weight", (128, 128)), ("k_proj.weight", (128, 128)), ("v_proj.weight", (128, 128)),
                         ("merge.weight", (128, 128)), ("mlp.0.weight", (256, 256)), ("mlp.2.weight", (128, 256))):
            bound = (6.0 / (shp[0] + shp[1])) ** 0.5
            sd[key] = ((torch.rand(shp, generator=gen) * 2 - 1) * bound).to(torch.bfloat16).float()
        for key in ("norm1", "norm2"):
            sd[key + ".weight"] = 1 + 0.2 * torch.randn(128, generator=gen)
            sd[key + ".bias"] = 0.1 * torch.randn(128, generator=gen)
        layers.append(sd)
    gain = float(rng.choice([0.3, 1.0, 3.0]))
    f0 = (gain * torch.randn(m, 25, 128, generator=gen)).to(torch.bfloat16).float()
    f1 = (gain * torch.randn(m, 25, 128, generator=gen)).to(torch.bfloat16).float()
    d0, d1 = f0.to(dev, torch.bfloat16), f1.to(dev, torch.bfloat16)
    ops.fine_transformer(d0, d1, torch.cat([ops.pack_fine_layer(sd, dev) for sd in layers]), names)
    torch.cuda.synchronize()
    w0, w1 = O.fine_transformer(f0, f1, layers, names)
    for got, want in ((d0.cpu(), w0), (d1.cpu(), w1)):
        rms, mx = rel(got, want)
        if not (bool(torch.isfinite(got.float()).all()) and rms < 1.5e-2 + 4e-3 * len(names) and mx < 8e-2 + 2e-2 * len(names)):
            bad += 1
            print(f"fine_tf case {case}: m={m} {names} gain={gain}: rms {rms:.4f} max {mx:.4f}")
            break
    # ---- (b) host pipeline vs device path
    n = int(rng.integers(1, 7))
    h0, w0_, h1, w1_ = (int(rng.integers(4, 28)) for _ in range(4))
    dtype = torch.bfloat16 if rng.random() < 0.6 else torch.float32
    thr = float(rng.choice([0.1, 0.2, 0.3]))
    chunk = int(rng.integers(1, n + 2))
    fc0, fc1 = synth.coarse_features(500 + case, n, h0 * w0_, h1 * w1_, 256, sigma=float(rng.uniform(0.7, 1.1)), dtype=dtype)
    ff0, _ = synth.fine_feature_maps(600 + case, n, h0 * 4, w0_ * 4, 128, dtype=dtype)
    _, ff1 = synth.fine_feature_maps(700 + case, n, h1 * 4, w1_ * 4, 128, dtype=dtype)
    desc = f"pipeline case {case}: n={n} {h0}x{w0_} vs {h1}x{w1_} {str(dtype)[6:]} thr={thr} chunk={chunk}"
    res = ops.match_pairs_device(fc0.to(dev), fc1.to(dev), ff0.to(dev), ff1.to(dev), (h0 * 8, w0_ * 8), (h0, w0_), (h1, w1_), thr=thr)
    mt = res.total()
    out = driver.match_pairs_host(fc0, fc1, ff0, ff1, (h0 * 8, w0_ * 8), (h0, w0_), (h1, w1_), chunk_pairs=chunk, device=0, thr=thr)
    cat = driver.flatten_slots(out)
    okp = int(out["counts"].sum()) == mt and all(torch.equal(cat[k], res[k][:mt].cpu()) for k in
                                                 ("b_ids", "i_ids", "j_ids", "mconf", "mkpts0_f", "mkpts1_f"))
    if not okp:
        bad += 1
        print(desc, "-> differs", int(out["counts"].sum()), mt)
print(f"fuzz_fine: {cases} cases of each, {bad} failures, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
