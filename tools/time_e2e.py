"""End-to-end host pipeline (pope_pipeline_run: pinned host buffers in, pinned host buffers out) on bench.py's workload,
over chunk sizes and the POPE_PIPELINE_F1 modes.  Wall clock around the call (it returns
after every copy has landed).  usage: python tools/time_e2e.py [pairs] [--chunks 8,16,32] [--modes union,windows]"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pope_b200 import driver, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("pairs", nargs="?", type=int, default=64)
    ap.add_argument("--chunks", default="8,16,32")
    ap.add_argument("--modes", default="union")
    ap.add_argument("--reps", type=int, default=4)
    a = ap.parse_args()
    n, hc, wc = a.pairs, 60, 80
    L = hc * wc
    f0, f1 = synth.coarse_features(1234, n, L, L, 256, dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(4321)
    driver.bind_host_to_gpu(0)
    h_f0, h_f1 = f0.pin_memory(), f1.pin_memory()
    h_ff0 = torch.randn(n, hc * 4, wc * 4, 128, generator=g).to(torch.bfloat16).pin_memory()
    h_ff1 = torch.randn(n, hc * 4, wc * 4, 128, generator=g).to(torch.bfloat16).pin_memory()
    rows = []
    for chunk in [int(c) for c in a.chunks.split(",")]:
        pl = driver.Pipeline(torch.bfloat16, chunk, (480, 640), (hc, wc), (hc, wc), device=0)
        out = pl.alloc_outputs(n)
        for mode in a.modes.split(","):
            os.environ["POPE_PIPELINE_F1"] = mode
            pl.run(h_f0, h_f1, h_ff0, h_ff1, out)
            t0 = time.perf_counter()
            for _ in range(a.reps):
                pl.run(h_f0, h_f1, h_ff0, h_ff1, out)
            dt = (time.perf_counter() - t0) / a.reps
            rows.append({"chunk": chunk, "mode": pl.last_f1_mode, "ms": round(1e3 * dt, 3),
                         "pairs_per_s": round(n / dt, 1), "h2d_gb": round(pl.last_h2d_bytes / 1e9, 4),
                         "h2d_gbs": round(pl.last_h2d_bytes / dt / 1e9, 2), "matches": int(out["counts"].sum())})
            print(json.dumps(rows[-1]), flush=True)
        pl.close()


if __name__ == "__main__":
    main()
