"""Padded batches (the -1e9 fill of coarse_matching.py:115-118 and `mask_border_with_padding`
:28-43).  CPU: the oracle restatement against the fixture produced by the UNMODIFIED reference (oracle/gen_golden_masked.py).
GPU: the drop-in CoarseMatching (CUDA path on the valid cells of every pair) against the same fixture and the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import pope_oracle as O
from oracle.gen_golden_masked import CASE, inputs, masks
from tests.parity_utils import compare_match_lists

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "coarse_masked.npz")


def _hw_i(case):
    return (case["hw0_c"][0] * 8, case["hw0_c"][1] * 8)


def test_fixture_matches_its_generator():
    assert json.loads(str(np.load(GOLD)["meta"])) == json.loads(json.dumps(CASE))


def test_oracle_masked_matches_reference_fixture():
    g = np.load(GOLD)
    f0, f1 = inputs(CASE)
    m0, m1 = masks(CASE)
    o = O.coarse_match_masked(f0, f1, _hw_i(CASE), CASE["hw0_c"], CASE["hw1_c"], m0, m1)
    for k in ("b_ids", "i_ids", "j_ids", "mkpts0_c", "mkpts1_c"):
        assert np.array_equal(o[k].numpy(), g[k]), k
    assert np.allclose(o["mconf"].numpy(), g["mconf"], rtol=1e-6, atol=0)
    assert len(g["b_ids"]) > 50
    # no match touches an invalid cell or the border of a valid rectangle
    for b, i, j in zip(g["b_ids"], g["i_ids"], g["j_ids"]):
        (h0, w0), (h1, w1) = CASE["valid0"][b], CASE["valid1"][b]
        y0, x0, y1, x1 = i // CASE["hw0_c"][1], i % CASE["hw0_c"][1], j // CASE["hw1_c"][1], j % CASE["hw1_c"][1]
        assert 2 <= y0 < h0 - 2 and 2 <= x0 < w0 - 2 and 2 <= y1 < h1 - 2 and 2 <= x1 < w1 - 2


def test_oracle_masked_equals_unmasked_oracle_on_the_valid_rectangle():
    """Property behind the CUDA implementation: masking = running the plain path on the valid cells only."""
    f0, f1 = inputs(CASE)
    m0, m1 = masks(CASE)
    o = O.coarse_match_masked(f0, f1, _hw_i(CASE), CASE["hw0_c"], CASE["hw1_c"], m0, m1)
    W0, W1 = CASE["hw0_c"][1], CASE["hw1_c"][1]
    for b in range(CASE["n"]):
        (h0, w0), (h1, w1) = CASE["valid0"][b], CASE["valid1"][b]
        v0, v1 = torch.nonzero(m0[b].reshape(-1)).reshape(-1), torch.nonzero(m1[b].reshape(-1)).reshape(-1)
        p = O.coarse_match(f0[b, v0][None], f1[b, v1][None], (h0 * 8, w0 * 8), (h0, w0), (h1, w1))
        sel = o["b_ids"] == b
        assert torch.equal(v0[p["i_ids"]], o["i_ids"][sel]) and torch.equal(v1[p["j_ids"]], o["j_ids"][sel])
        assert torch.allclose(p["mconf"], o["mconf"][sel], rtol=1e-6, atol=0)
        assert W0 >= w0 and W1 >= w1


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("mask_in_data", [True, False])
def test_cuda_coarse_matching_with_padding_masks(dtype, mask_in_data):
    import pope_b200
    dev = torch.device("cuda:0")
    g = np.load(GOLD)
    f0, f1 = inputs(CASE)
    if dtype == torch.bfloat16:
        f0, f1 = f0.to(dtype).float(), f1.to(dtype).float()            # the oracle sees the same rounded values
    m0, m1 = masks(CASE)
    hw0_c, hw1_c = CASE["hw0_c"], CASE["hw1_c"]
    cm = pope_b200.CoarseMatching(pope_b200.make_default_cfg()["match_coarse"]).eval()
    cm.materialize_conf_matrix = dtype == torch.float32            # the debug output, checked below
    data = {"hw0_i": torch.Size(_hw_i(CASE)), "hw1_i": torch.Size((hw1_c[0] * 8, hw1_c[1] * 8)),
            "hw0_c": torch.Size(hw0_c), "hw1_c": torch.Size(hw1_c)}
    if mask_in_data:
        data.update(mask0=m0.to(dev), mask1=m1.to(dev))
    cm(f0.to(dev, dtype), f1.to(dev, dtype), data, mask_c0=m0.flatten(-2).to(dev), mask_c1=m1.flatten(-2).to(dev))
    if mask_in_data:
        want = O.coarse_match_masked(f0, f1, _hw_i(CASE), hw0_c, hw1_c, m0, m1)
    else:       # masks passed to forward() only: -1e9 fill, but the plain border of the full grids (coarse_matching.py:180-181)
        full = O.coarse_match_masked(f0, f1, _hw_i(CASE), hw0_c, hw1_c, m0, m1, border_rm=0)
        k0, k1 = O._interior(hw0_c[0], hw0_c[1], 2), O._interior(hw1_c[0], hw1_c[1], 2)
        keep = k0[full["i_ids"]] & k1[full["j_ids"]]
        want = {k: full[k][keep] for k in ("b_ids", "i_ids", "j_ids", "mconf", "mkpts0_c", "mkpts1_c")}
        want["conf_matrix"] = full["conf_matrix"]
    conf = want["conf_matrix"]
    mg = dict(conf_rowmax=conf.topk(2, dim=2)[0][..., 0].numpy(), conf_row2nd=conf.topk(2, dim=2)[0][..., 1].numpy(),
              conf_colmax=conf.topk(2, dim=1)[0][:, 0].numpy(), conf_col2nd=conf.topk(2, dim=1)[0][:, 1].numpy())
    got = {k: data[k].cpu() for k in ("b_ids", "i_ids", "j_ids", "mconf", "mkpts0_c", "mkpts1_c")}
    same, near, bad = compare_match_lists(got, want, mg)
    assert not bad and len(near) <= 1 and same > 40, (same, near, bad)
    if not near:
        assert torch.allclose(got["mconf"], want["mconf"], rtol=1e-2 if dtype == torch.bfloat16 else 1e-4, atol=0)
        assert torch.equal(got["mkpts0_c"], want["mkpts0_c"].float()) and torch.equal(got["mkpts1_c"], want["mkpts1_c"].float())
        if mask_in_data and dtype == torch.float32:
            for k in ("b_ids", "i_ids", "j_ids"):
                assert np.array_equal(got[k].numpy(), g[k]), k
    assert data["gt_mask"].dtype == torch.bool and torch.equal(data["m_bids"], data["b_ids"])
    if dtype == torch.float32:
        assert torch.allclose(data["conf_matrix"].cpu(), conf, rtol=1e-4, atol=1e-7)
    else:
        assert "conf_matrix" not in data
