"""The oracle (oracle/pope_oracle.py) against the fixtures produced by the unmodified reference
(oracle/gen_golden.py -> tests/golden/).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import pope_oracle as O
from oracle.gen_golden import COARSE_CASES, coarse_inputs
from pope_b200 import synth


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


@pytest.mark.parametrize("name", list(COARSE_CASES))
def test_coarse_match_matches_reference(golden_dir, name):
    case = COARSE_CASES[name]
    g = _load(golden_dir, name)
    assert json.loads(str(g["meta"])) == json.loads(json.dumps(case))
    f0, f1 = coarse_inputs(case)
    hw0_c, hw1_c = case["hw0_c"], case["hw1_c"]
    out = O.coarse_match(f0, f1, (hw0_c[0] * 8, hw0_c[1] * 8), hw0_c, hw1_c)
    for k in ("b_ids", "i_ids", "j_ids", "m_bids"):
        assert out[k].dtype == torch.int64
        np.testing.assert_array_equal(out[k].numpy(), g[k], err_msg=k)
    np.testing.assert_array_equal(out["gt_mask"].numpy(), g["gt_mask"])
    for k in ("mconf", "mkpts0_c", "mkpts1_c"):
        assert out[k].dtype == torch.float32
        np.testing.assert_allclose(out[k].numpy(), g[k], rtol=1e-6, atol=0, err_msg=k)


def test_coarse_match_chunking_is_transparent():
    f0, f1 = synth.coarse_features(3, 5, 96, 80, 64, sigma=0.8)
    a = O.coarse_match(f0, f1, (64, 96), (8, 12), (8, 10), chunk=5)
    b = O.coarse_match(f0, f1, (64, 96), (8, 12), (8, 10), chunk=2)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_fine_windows_match_reference(golden_dir):
    g = _load(golden_dir, "fine_preprocess")
    meta = json.loads(str(g["meta"]))
    case = COARSE_CASES[meta["coarse_case"]]
    c = _load(golden_dir, meta["coarse_case"])
    b_ids, i_ids, j_ids = (torch.from_numpy(c[k]) for k in ("b_ids", "i_ids", "j_ids"))
    hw = case["hw0_c"]
    ff0, ff1 = synth.fine_feature_maps(meta["fine_seed"], case["n"], hw[0] * 4, hw[1] * 4, 128, channels_last=False)
    for feat, ids, key in ((ff0, i_ids, "win0"), (ff1, j_ids, "win1")):
        w = O.fine_windows(feat, b_ids, ids)
        wd = O.fine_windows_direct(feat, b_ids, ids, hw[1])
        assert torch.equal(w, wd)
        np.testing.assert_array_equal(w[: meta["kept"]].numpy(), g[key])
        np.testing.assert_allclose(w.sum((1, 2)).numpy(), g[key + "_sum"], rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("name", ["fine_match_soft", "fine_match_peaked"])
def test_fine_match_matches_reference(golden_dir, name):
    g = _load(golden_dir, name)
    meta = json.loads(str(g["meta"]))
    a, b = synth.fine_windows(meta["seed"], meta["M"], 25, 128, gain=meta["gain"])
    out = O.fine_match(a, b, torch.from_numpy(g["mkpts0_c"]), torch.from_numpy(g["mkpts1_c"]), scale=2.0)
    np.testing.assert_allclose(out["expec_f"][:, :2].numpy(), g["expec_f"][:, :2], rtol=1e-5, atol=1e-6)
    # std = sum sqrt(clamp(E[g^2]-E[g]^2)) cancels catastrophically for peaked heatmaps (fine_matching.py:53-54)
    np.testing.assert_allclose(out["expec_f"][:, 2].numpy(), g["expec_f"][:, 2], rtol=1e-4, atol=2e-3)
    np.testing.assert_array_equal(out["mkpts0_f"].numpy(), g["mkpts0_f"])
    np.testing.assert_allclose(out["mkpts1_f"].numpy(), g["mkpts1_f"], rtol=1e-6, atol=1e-5)


def test_fine_match_empty():
    out = O.fine_match(torch.empty(0, 25, 128), torch.empty(0, 25, 128), torch.empty(0, 2), torch.empty(0, 2), 2.0)
    assert out["expec_f"].shape == (0, 3) and out["mkpts0_f"].shape == (0, 2) and out["mkpts1_f"].shape == (0, 2)


# ---- known-answer tests; the reference has none of its own -----------------------

def test_kat_permutation_gives_interior_matches():
    h, w, C = 8, 10, 64
    L = h * w
    g = torch.Generator().manual_seed(0)
    f0 = 4.0 * torch.randn(1, L, C, generator=g)
    perm = torch.randperm(L, generator=g)
    f1 = torch.empty_like(f0)
    f1[0, perm] = f0[0]                      # cell i of image 0 <-> cell perm[i] of image 1
    out = O.coarse_match(f0, f1, (h * 8, w * 8), (h, w), (h, w))
    keep = O._interior(h, w, 2)
    want = [(i, int(perm[i])) for i in range(L) if keep[i] and keep[perm[i]]]
    got = list(zip(out["i_ids"].tolist(), out["j_ids"].tolist()))
    assert got == want
    assert out["mconf"].min() > 0.999


def test_kat_border_cell_vetoes_interior_candidate():
    h, w, C = 8, 8, 32
    L = h * w
    f0 = torch.zeros(1, L, C)
    f1 = torch.zeros(1, L, C)
    i = 3 * w + 3                 # interior query cell
    j_in = 4 * w + 4              # interior reference cell (weaker)
    j_bd = 0 * w + 5              # border reference cell (stronger)
    f0[0, i, 0] = 16.0
    f1[0, j_in, 0] = 16.0 * 0.60
    f1[0, j_bd, 0] = 16.0 * 0.62
    out = O.coarse_match(f0, f1, (64, 64), (h, w), (h, w), thr=0.05)
    assert out["i_ids"].numel() == 0          # the border cell holds the row maximum and is then removed
    f1[0, j_bd, 0] = 16.0 * 0.55              # now the interior cell wins
    out = O.coarse_match(f0, f1, (64, 64), (h, w), (h, w), thr=0.05)
    assert out["i_ids"].tolist() == [i] and out["j_ids"].tolist() == [j_in]


def test_kat_fine_one_hot_and_uniform():
    M, WW, C = 25, 25, 128
    w0 = torch.zeros(M, WW, C)
    w1 = torch.zeros(M, WW, C)
    w0[:, 12, 0] = 100.0
    for r in range(25):
        w1[r, r, 0] = 100.0        # match r has its peak at window position r
    out = O.fine_match(w0, w1, torch.zeros(M, 2), torch.zeros(M, 2), 2.0)
    lin = torch.linspace(-1, 1, 5)
    want = torch.stack([lin.repeat(5), lin.repeat_interleave(5)], 1)
    assert torch.allclose(out["expec_f"][:, :2], want, atol=1e-6)
    assert torch.allclose(out["expec_f"][:, 2], torch.full((M,), 2e-5), atol=1e-3)
    assert torch.allclose(out["mkpts1_f"], want * 4.0, atol=1e-5)
    out = O.fine_match(torch.zeros(3, WW, C), torch.zeros(3, WW, C), torch.zeros(3, 2), torch.zeros(3, 2), 2.0)
    assert torch.allclose(out["expec_f"][:, :2], torch.zeros(3, 2), atol=1e-7)
    assert torch.allclose(out["expec_f"][:, 2], torch.full((3,), 2 * 0.5 ** 0.5), atol=1e-6)


def test_kat_running_topk_slot_semantics():
    s, i = O.running_topk([0.5, -0.2, 0.3, 0.9, 0.1, 0.4], 3)
    # 0.5->slot0, -0.2 skipped, 0.3->slot1, 0.9->slot2, 0.1 skipped (not > any slot), 0.4 replaces the min (0.3, slot1)
    assert s == [0.5, 0.4, 0.9] and i == [0, 5, 3]
    s, i = O.running_topk([-1.0, -0.5], 3)
    assert i == [-1, -1, -1]
    q, refs = synth.retrieval_tokens(7, 64, 384)
    sc = O.cosine_scores(q, refs)
    s, i = O.running_topk(sc.tolist(), 3)
    assert sorted(i) == sorted(torch.topk(sc, 3)[1].tolist())
