"""Generates tests/golden/coarse_masked.npz: CoarseMatching of the UNMODIFIED reference on padded batches (padding masks as
the reference's MegaDepth loader produces them: the valid area of every image is a top-left rectangle of the coarse grid).
Run in the build container (needs /root/reference):   python oracle/gen_golden_masked.py"""
import copy
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim                      # noqa: E402
from pope_b200 import synth                      # noqa: E402

CASE = dict(seed=31, n=4, hw0_c=(14, 18), hw1_c=(12, 16), C=64, sigma=0.8,
            valid0=[(14, 18), (10, 18), (14, 11), (7, 9)], valid1=[(12, 16), (12, 9), (8, 16), (12, 16)])


def masks(case):
    m0 = torch.zeros(case["n"], *case["hw0_c"], dtype=torch.bool)
    m1 = torch.zeros(case["n"], *case["hw1_c"], dtype=torch.bool)
    for b, ((h0, w0), (h1, w1)) in enumerate(zip(case["valid0"], case["valid1"])):
        m0[b, :h0, :w0] = True
        m1[b, :h1, :w1] = True
    return m0, m1


def inputs(case):
    L, S = case["hw0_c"][0] * case["hw0_c"][1], case["hw1_c"][0] * case["hw1_c"][1]
    return synth.coarse_features(case["seed"], case["n"], L, S, case["C"], sigma=case["sigma"], planted=0.9)


def main():
    ref = ref_shim.import_reference()
    from src.matcher.utils.coarse_matching import CoarseMatching
    torch.set_grad_enabled(False)
    cm = CoarseMatching(copy.deepcopy(ref.default_cfg)["match_coarse"]).eval()
    f0, f1 = inputs(CASE)
    m0, m1 = masks(CASE)
    hw0_c, hw1_c = CASE["hw0_c"], CASE["hw1_c"]
    data = {"hw0_i": torch.Size([hw0_c[0] * 8, hw0_c[1] * 8]), "hw1_i": torch.Size([hw1_c[0] * 8, hw1_c[1] * 8]),
            "hw0_c": torch.Size(hw0_c), "hw1_c": torch.Size(hw1_c), "mask0": m0, "mask1": m1}
    cm(f0, f1, data, mask_c0=m0.flatten(-2), mask_c1=m1.flatten(-2))          # as src/matcher/matcher.py:62-71 calls it
    conf = data["conf_matrix"]
    top2_row, top2_col = conf.topk(2, dim=2)[0], conf.topk(2, dim=1)[0]
    out = os.path.join(ROOT, "tests", "golden", "coarse_masked.npz")
    np.savez_compressed(out, meta=json.dumps(CASE), b_ids=data["b_ids"].numpy(), i_ids=data["i_ids"].numpy(),
                        j_ids=data["j_ids"].numpy(), mconf=data["mconf"].numpy(), mkpts0_c=data["mkpts0_c"].numpy(),
                        mkpts1_c=data["mkpts1_c"].numpy(), conf_rowmax=top2_row[..., 0].numpy(),
                        conf_row2nd=top2_row[..., 1].numpy(), conf_colmax=top2_col[:, 0].numpy(),
                        conf_col2nd=top2_col[:, 1].numpy())
    per = np.bincount(data["b_ids"].numpy(), minlength=CASE["n"])
    print("coarse_masked: M per pair", per.tolist(), "->", out)


if __name__ == "__main__":
    main()
