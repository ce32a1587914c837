"""Parity of the CUDA path (through the C ABI) against the reference-generated golden fixtures and the oracle.
Run on the B200 box:  python -m pytest tests -m gpu"""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import pope_oracle as O
from oracle.gen_golden import COARSE_CASES, coarse_inputs
from pope_b200 import _lib, ops, synth
from tests.parity_utils import assert_sorted_by_pair_and_row, compare_match_lists, oracle_with_margins

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
IMPLS = {"simt": _lib.COARSE_SIMT, "tcgen05": _lib.COARSE_TCGEN05}
_ORACLE_CACHE = {}


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def _tc_ok(C, dtype, L=4800, S=4800):
    """bf16 features run the tcgen05 kernels directly; fp32 features run them through the three-way bf16 split
    (same shapes; ops.coarse_match allocates the larger workspace that path needs)."""
    return dtype in (torch.bfloat16, torch.float32) and _lib.tcgen05_available(L, S, C)


def _need_tc(impl, C=256, L=4800, S=4800):
    if impl == "tcgen05" and not _lib.tcgen05_available(L, S, C):
        pytest.skip("tcgen05 kernels do not cover this shape")


def _run_coarse(f0, f1, hw0_c, hw1_c, impl, dtype, **kw):
    res = ops.coarse_match(f0.to(DEV, dtype), f1.to(DEV, dtype), hw0_c, hw1_c, 8.0, impl=impl, **kw)
    # POPE_FLAG_ROBUST_PATH is informational (the single-sweep kernel handed the batch to the online-softmax kernels)
    assert res.flags() & ~_lib.FLAG_ROBUST_PATH == 0
    out = {k: v.cpu() for k, v in res.sliced().items()}
    out["_flags"] = res.flags()
    counts = res["counts"][: f0.shape[0]].cpu()
    assert int(counts.sum()) == out["b_ids"].numel()
    assert torch.equal(torch.bincount(out["b_ids"], minlength=f0.shape[0]).to(torch.int32), counts)
    return out


@pytest.mark.parametrize("impl", list(IMPLS))
@pytest.mark.parametrize("name", list(COARSE_CASES))
def test_coarse_golden(golden_dir, name, impl):
    case = COARSE_CASES[name]
    g = _load(golden_dir, name)
    f0, f1 = coarse_inputs(case)            # fp32 values; for the bf16 case they are exactly bf16-representable
    dtypes = [torch.bfloat16] if case["kind"] == "bf16" else [torch.float32]
    for dtype in dtypes:
        if impl == "tcgen05" and not _tc_ok(case["C"], dtype, case["hw0_c"][0] * case["hw0_c"][1],
                                            case["hw1_c"][0] * case["hw1_c"][1]):
            pytest.skip("tcgen05 path takes features with C in {64,128,192,256}")
        out = _run_coarse(f0, f1, case["hw0_c"], case["hw1_c"], IMPLS[impl], dtype)
        want = {k: torch.from_numpy(g[k]) for k in ("b_ids", "i_ids", "j_ids", "mconf", "mkpts0_c", "mkpts1_c")}
        same, near, bad = compare_match_lists(out, want, g)
        assert not bad, f"{name}/{impl}: index mismatches that are not near-ties: {bad[:5]}"
        assert len(near) <= max(2, want["b_ids"].numel() // 200), near
        if not near:
            for k in ("b_ids", "i_ids", "j_ids"):
                assert out[k].dtype == torch.int64 and torch.equal(out[k], want[k]), k
            tol = 1e-2 if dtype == torch.bfloat16 else 1e-4
            if case["kind"] == "hard":
                # |S| reaches several hundred here: one fp32 ulp of S is ~3e-5 and conf = exp(2S - lse_r - lse_c)
                # inherits it on BOTH sides (the reference's fp32 einsum included), so 1e-4 is not attainable
                tol = 5e-4
            np.testing.assert_allclose(out["mconf"].numpy(), g["mconf"], rtol=tol, atol=0)
            np.testing.assert_array_equal(out["mkpts0_c"].numpy(), g["mkpts0_c"])
            np.testing.assert_array_equal(out["mkpts1_c"].numpy(), g["mkpts1_c"])
            assert out["gt_mask"].dtype == torch.bool and not out["gt_mask"].any()
        L = case["hw0_c"][0] * case["hw0_c"][1]
        if out["b_ids"].numel() > 1:
            assert_sorted_by_pair_and_row(out["b_ids"], out["i_ids"], L)


@pytest.mark.parametrize("impl", list(IMPLS))
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_coarse_full_size_vs_oracle(impl, dtype):
    """BASELINE configs[1] shape (480x640 -> 60x80 tokens, d=256), 3 pairs, against the oracle on the same
    (bf16-rounded) values."""
    _need_tc(impl)
    f0, f1 = synth.coarse_features(41, 3, 4800, 4800, 256, dtype=dtype)
    want, mg = oracle_with_margins(f0.float(), f1.float(), (480, 640), (60, 80), (60, 80))
    out = _run_coarse(f0, f1, (60, 80), (60, 80), IMPLS[impl], dtype)
    same, near, bad = compare_match_lists(out, want, mg)
    assert not bad, bad[:5]
    assert len(near) <= 8, near
    assert same > 6000
    if not near:
        tol = 1e-2 if dtype == torch.bfloat16 else 1e-4
        assert torch.allclose(out["mconf"], want["mconf"], rtol=tol, atol=0)
        assert torch.equal(out["mkpts1_c"], want["mkpts1_c"])


@pytest.mark.parametrize("impl", list(IMPLS))
def test_coarse_hard_set_bf16(impl):
    """Exact duplicates, near-duplicates and x50 rows (|S| in the hundreds) on bf16-rounded inputs."""
    _need_tc(impl, 256, 40 * 48, 36 * 56)
    f0, f1 = synth.hard_coarse_features(43, 2, 40 * 48, 36 * 56, 256, sigma=0.9, dtype=torch.bfloat16)
    want, mg = oracle_with_margins(f0.float(), f1.float(), (320, 384), (40, 48), (36, 56))
    out = _run_coarse(f0, f1, (40, 48), (36, 56), IMPLS[impl], torch.bfloat16)
    same, near, bad = compare_match_lists(out, want, mg)
    assert not bad, bad[:5]
    assert same > 100
    if impl == "tcgen05":
        # a x50 row sits 500+ log2 units above the other 31 rows of its 32-row group: more than one warp-uniform fp32 shift
        # can hold, so the single sweep must have detected it and handed these pairs to the online-softmax launch
        assert out["_flags"] & _lib.FLAG_ROBUST_PATH


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_coarse_large_norm_stays_on_the_single_sweep(dtype):
    """Token norm ~ 49 (what the coarse transformer emits): |S| log2(e) reaches ~135, far outside what
    unshifted fp32 exponentials hold.  The lazily shifted single sweep must stay on the fast path (no flag) and agree with
    the oracle; so must a batch that mixes weak pairs (sigma 0.25: every similarity below 1) with such strong ones."""
    h0, w0, h1, w1 = 40, 48, 36, 56
    _need_tc("tcgen05", 256, h0 * w0, h1 * w1)
    fa0, fa1 = synth.coarse_features(91, 2, h0 * w0, h1 * w1, 256, sigma=49.0 / 16.0, dtype=dtype)
    fb0, fb1 = synth.coarse_features(92, 1, h0 * w0, h1 * w1, 256, sigma=0.25, noise=0.05, dtype=dtype)
    f0, f1 = torch.cat([fa0[:1], fb0, fa0[1:]]), torch.cat([fa1[:1], fb1, fa1[1:]])
    want, mg = oracle_with_margins(f0.float(), f1.float(), (h0 * 8, w0 * 8), (h0, w0), (h1, w1))
    out = _run_coarse(f0, f1, (h0, w0), (h1, w1), _lib.COARSE_TCGEN05, dtype)
    assert out["_flags"] == 0
    same, near, bad = compare_match_lists(out, want, mg)
    assert not bad, bad[:5]
    assert same > 1500 and len(near) <= 6
    if not near:
        assert torch.allclose(out["mconf"], want["mconf"], rtol=1e-2 if dtype == torch.bfloat16 else 1e-4, atol=0)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_coarse_shift_rises_along_the_sweep(dtype):
    """Columns whose norm grows geometrically from the first to the last tile (x1 ... x2.5: the planted similarities climb
    from ~20 to ~100 log2 units along a row), so that an epilogue warp has to RAISE its shift several times while it sweeps
    a row block (a raise is due whenever a cell exceeds the largest one so far by 14 units; the running row sums are
    rescaled each time), and rows scaled the other way so that the 32-row groups start from different shifts and the column
    merge has to combine partial sums taken under different shifts.  The spread inside a 32-row group (~95 units) is within
    what one shift holds (186), so the pair must stay on the single sweep; results equal the oracle."""
    h0, w0, h1, w1 = 40, 48, 36, 56
    L, S = h0 * w0, h1 * w1
    _need_tc("tcgen05", 256, L, S)
    f0, f1 = synth.coarse_features(95, 2, L, S, 256, sigma=1.2, noise=0.2, dtype=torch.float32)
    col_gain = torch.logspace(0, math.log10(2.5), S)                    # image-1 cells: x1 .. x2.5 in sweep order
    row_gain = torch.logspace(math.log10(2.0), 0, L)                    # image-0 cells: x2 .. x1
    f1 = (f1 * col_gain[None, :, None]).to(dtype)
    f0 = (f0 * row_gain[None, :, None]).to(dtype)
    want, mg = oracle_with_margins(f0.float(), f1.float(), (h0 * 8, w0 * 8), (h0, w0), (h1, w1))
    out = _run_coarse(f0, f1, (h0, w0), (h1, w1), _lib.COARSE_TCGEN05, dtype)
    assert out["_flags"] == 0
    same, near, bad = compare_match_lists(out, want, mg)
    assert not bad, bad[:5]
    assert same > 500 and len(near) <= 6
    if not near:
        assert torch.allclose(out["mconf"], want["mconf"], rtol=1e-2 if dtype == torch.bfloat16 else 1e-4, atol=0)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
def test_coarse_split_sweep_equals_single_launch(dtype):
    """When the static unit schedule's last round is part empty the single sweep runs as two launches (head + tail) with the
    head's column merge beside the tail (csrc/coarse_tc.cu::coarse_tc_run, and coarse_tc_split_run for fp32 features).  Same
    arithmetic, so the results must equal the one-launch form (developer bit 10) bit for bit -- also when a flagged pair
    sits in the head or in the tail."""
    h, w = 40, 48                                   # 1920 cells: 8 units per pair; 24 pairs = 192 units = 2 rounds of 74 + 44
    n = 24
    assert _lib.single_sweep_is_split(n, h * w)
    _need_tc("tcgen05", 256, h * w, h * w)
    f0, f1 = synth.coarse_features(97, n, h * w, h * w, 256, sigma=1.1, dtype=dtype)
    fh0, fh1 = synth.hard_coarse_features(98, 2, h * w, h * w, 256, sigma=0.9, dtype=dtype)
    for variant in ("plain", "flagged"):
        if variant == "flagged":                    # pair 3 lies in the head (18 pairs), pair 23 in the tail
            f0[3], f1[3], f0[23], f1[23] = fh0[0], fh1[0], fh0[1], fh1[1]
        outs = []
        for knob in ("0", "1024"):
            os.environ["POPE_TC_DEBUG"] = knob
            try:
                outs.append(_run_coarse(f0, f1, (h, w), (h, w), _lib.COARSE_TCGEN05, dtype))
            finally:
                del os.environ["POPE_TC_DEBUG"]
        assert bool(outs[0]["_flags"] & _lib.FLAG_ROBUST_PATH) == (variant == "flagged")
        assert outs[0]["b_ids"].numel() > 10000
        for k in outs[0]:
            if k != "_flags":
                assert torch.equal(outs[0][k], outs[1][k]), (variant, k)
        assert outs[0]["_flags"] == outs[1]["_flags"]


def test_coarse_mixed_batch_only_flagged_pairs_take_the_robust_launch():
    """A batch of ordinary pairs with one hard-set pair in the middle: the flag is raised, the result of every pair equals
    what the pair gives when it is run alone (the gated launch redoes the flagged pair only; the others keep the
    single-sweep results bit for bit)."""
    h0, w0, h1, w1 = 40, 48, 36, 56
    _need_tc("tcgen05", 256, h0 * w0, h1 * w1)
    fa0, fa1 = synth.coarse_features(93, 2, h0 * w0, h1 * w1, 256, sigma=0.9, dtype=torch.bfloat16)
    fh0, fh1 = synth.hard_coarse_features(94, 1, h0 * w0, h1 * w1, 256, sigma=0.9, dtype=torch.bfloat16)
    f0, f1 = torch.cat([fa0[:1], fh0, fa0[1:]]), torch.cat([fa1[:1], fh1, fa1[1:]])
    out = _run_coarse(f0, f1, (h0, w0), (h1, w1), _lib.COARSE_TCGEN05, torch.bfloat16)
    assert out["_flags"] & _lib.FLAG_ROBUST_PATH
    for b in range(3):
        solo = _run_coarse(f0[b:b + 1], f1[b:b + 1], (h0, w0), (h1, w1), _lib.COARSE_TCGEN05, torch.bfloat16)
        assert bool(solo["_flags"] & _lib.FLAG_ROBUST_PATH) == (b == 1)
        sel = out["b_ids"] == b
        assert int(sel.sum()) == solo["b_ids"].numel() > 100
        for k in ("i_ids", "j_ids", "mconf", "mkpts1_c"):
            assert torch.equal(out[k][sel], solo[k]), (b, k)


def test_coarse_fp32_on_tensor_cores_vs_fp32_fma():
    """fp32 features through the tcgen05 split path (a = a1 + a2 + a3 in bf16, six products, fp32 accumulation) against
    the fp32-FMA kernels on the same inputs: identical match lists, confidences to a few fp32 ulps of the logits; on
    out-of-range data (rows x50) the split path must hand over to the fp32-FMA kernels (flag) and then be bit-identical."""
    L, S = 50 * 70, 44 * 60
    _need_tc("tcgen05", 256, L, S)
    f0, f1 = synth.coarse_features(48, 3, L, S, 256)
    tc = _run_coarse(f0, f1, (50, 70), (44, 60), IMPLS["tcgen05"], torch.float32)
    fma = _run_coarse(f0, f1, (50, 70), (44, 60), IMPLS["simt"], torch.float32)
    assert tc["_flags"] == 0 and tc["b_ids"].numel() > 3000
    for k in ("b_ids", "i_ids", "j_ids"):
        assert torch.equal(tc[k], fma[k]), k
    assert torch.allclose(tc["mconf"], fma["mconf"], rtol=3e-5, atol=0)
    h0, h1 = synth.hard_coarse_features(49, 2, 40 * 48, 36 * 56, 256, sigma=0.9)
    tc = _run_coarse(h0, h1, (40, 48), (36, 56), IMPLS["tcgen05"], torch.float32)
    fma = _run_coarse(h0, h1, (40, 48), (36, 56), IMPLS["simt"], torch.float32)
    assert tc["_flags"] & _lib.FLAG_ROBUST_PATH
    for k in ("b_ids", "i_ids", "j_ids", "mconf"):
        assert torch.equal(tc[k], fma[k]), k


def test_coarse_single_sweep_vs_robust_path():
    """tcgen05: the single-sweep kernel (lazily shifted 2^(x - m), shuffle-reduced column sums) and the two-sweep online-softmax
    kernels it falls back to must produce the same match lists on in-range data (ragged L != S, 3 pairs)."""
    L, S = 50 * 70, 44 * 60
    _need_tc("tcgen05", 256, L, S)
    f0, f1 = synth.coarse_features(47, 3, L, S, 256, dtype=torch.bfloat16)
    fast = _run_coarse(f0, f1, (50, 70), (44, 60), IMPLS["tcgen05"], torch.bfloat16)
    assert fast["_flags"] == 0
    os.environ["POPE_TC_DEBUG"] = "16"          # developer knob: skip the single sweep
    try:
        slow = _run_coarse(f0, f1, (50, 70), (44, 60), IMPLS["tcgen05"], torch.bfloat16)
    finally:
        del os.environ["POPE_TC_DEBUG"]
    assert fast["b_ids"].numel() > 3000
    for k in ("b_ids", "i_ids", "j_ids"):
        assert torch.equal(fast[k], slow[k]), k
    assert torch.allclose(fast["mconf"], slow["mconf"], rtol=2e-5, atol=0)


@pytest.mark.parametrize("impl", list(IMPLS))
@pytest.mark.parametrize("thr", [0.1, 0.3])
def test_coarse_other_thresholds(impl, thr):
    """thr = 0.1 takes the tcgen05 three-sweep path (too many candidates per row for the list), 0.3 the two-sweep one."""
    _need_tc(impl, 256, 1200, 1024)
    f0, f1 = synth.coarse_features(45, 2, 1200, 1024, 256, sigma=0.85, dtype=torch.bfloat16)
    want, mg = oracle_with_margins(f0.float(), f1.float(), (240, 320), (30, 40), (32, 32), thr=thr)
    out = _run_coarse(f0, f1, (30, 40), (32, 32), IMPLS[impl], torch.bfloat16, thr=thr)
    same, near, bad = compare_match_lists(out, want, mg, thr=thr)
    assert not bad, bad[:5]
    assert len(near) <= 3 and same > 100
    if not near:
        assert torch.allclose(out["mconf"], want["mconf"], rtol=1e-2, atol=0)


@pytest.mark.parametrize("impl", list(IMPLS))
def test_coarse_permutation_property_full_size(impl):
    """Size-independent property at the benchmark size: f1 = P f0 with large norm -> the matches are exactly the
    interior cells with j = P(i), mconf -> 1; sorted by (b, i); every j used once."""
    _need_tc(impl)
    n, h, w, C = 4, 60, 80, 256
    L = h * w
    g = torch.Generator().manual_seed(5)
    f0 = (3.0 * torch.randn(n, L, C, generator=g)).to(torch.bfloat16)
    f1 = torch.empty_like(f0)
    perms = []
    for b in range(n):
        p = torch.randperm(L, generator=g)
        f1[b, p] = f0[b]
        perms.append(p)
    out = _run_coarse(f0, f1, (h, w), (h, w), IMPLS[impl], torch.bfloat16)
    assert out["_flags"] == 0        # |S| log2(e) ~ 130 on the diagonal: the lazily shifted single sweep holds it
    keep = O._interior(h, w, 2)
    want_b, want_i, want_j = [], [], []
    for b in range(n):
        i = torch.nonzero(keep & keep[perms[b]]).flatten()
        want_b.append(torch.full_like(i, b)); want_i.append(i); want_j.append(perms[b][i])
    assert torch.equal(out["b_ids"], torch.cat(want_b))
    assert torch.equal(out["i_ids"], torch.cat(want_i))
    assert torch.equal(out["j_ids"], torch.cat(want_j))
    assert float(out["mconf"].min()) > 0.999
    assert_sorted_by_pair_and_row(out["b_ids"], out["i_ids"], L)


def test_coarse_high_res_tcgen05_equals_fp32_fma():
    """BASELINE configs[3]: 960x1280 -> 19,200 coarse tokens per image.  The CPU oracle needs 7 GB for this shape, so
    the tensor-core path is held to the fp32-FMA kernels (themselves held to the oracle at the smaller sizes)."""
    _need_tc("tcgen05", 256, 19200, 19200)
    f0, f1 = synth.coarse_features(47, 1, 19200, 19200, 256, dtype=torch.bfloat16)
    a = _run_coarse(f0, f1, (120, 160), (120, 160), _lib.COARSE_SIMT, torch.bfloat16)
    b = _run_coarse(f0, f1, (120, 160), (120, 160), _lib.COARSE_TCGEN05, torch.bfloat16)
    assert a["b_ids"].numel() > 50
    for k in ("b_ids", "i_ids", "j_ids"):
        assert torch.equal(a[k], b[k]), k
    assert torch.allclose(a["mconf"], b["mconf"], rtol=1e-4, atol=0)
    assert torch.equal(a["mkpts1_c"], b["mkpts1_c"])


@pytest.mark.parametrize("impl", list(IMPLS))
def test_coarse_high_res_vs_oracle(impl):
    """BASELINE configs[3] against the ORACLE itself: one 960x1280 pair (19,200 coarse tokens per image; the reference's
    conf matrix is 1.47 GB and its op sequence peaks near 7 GB of host memory), bf16-rounded features,
    both kernel families.  Index sets bit-exact outside the near-tie / near-threshold classes, confidences within 1e-2."""
    _need_tc(impl, 256, 19200, 19200)
    f0, f1 = synth.coarse_features(53, 1, 19200, 19200, 256, dtype=torch.bfloat16)
    if "hi" not in _ORACLE_CACHE:          # ~1 minute of host time: once for both implementations
        _ORACLE_CACHE["hi"] = oracle_with_margins(f0.float(), f1.float(), (960, 1280), (120, 160), (120, 160))
    want, mg = _ORACLE_CACHE["hi"]
    out = _run_coarse(f0, f1, (120, 160), (120, 160), IMPLS[impl], torch.bfloat16)
    if impl == "tcgen05":
        assert out["_flags"] == 0
    same, near, bad = compare_match_lists(out, want, mg)
    assert not bad, bad[:5]
    assert want["b_ids"].numel() > 50 and same >= want["b_ids"].numel() - 4 and len(near) <= 4, (same, near)
    if not near:
        assert torch.allclose(out["mconf"], want["mconf"], rtol=1e-2, atol=0)
        assert torch.equal(out["mkpts1_c"], want["mkpts1_c"]) and torch.equal(out["mkpts0_c"], want["mkpts0_c"])
    assert_sorted_by_pair_and_row(out["b_ids"], out["i_ids"], 19200)


def test_coarse_edge_cases():
    # grid too small for the border -> no match, empty tensors of the right shapes/dtypes
    f0, f1 = synth.coarse_features(1, 1, 16, 16, 64, sigma=2.0)
    out = _run_coarse(f0, f1, (4, 4), (4, 4), _lib.COARSE_SIMT, torch.float32)
    assert out["b_ids"].shape == (0,) and out["mkpts0_c"].shape == (0, 2) and out["mconf"].dtype == torch.float32
    # border_rm = 0 keeps everything the mutual test keeps
    out0 = _run_coarse(f0, f1, (4, 4), (4, 4), _lib.COARSE_SIMT, torch.float32, border_rm=0)
    want = O.coarse_match(f0, f1, (32, 32), (4, 4), (4, 4), border_rm=0)
    assert torch.equal(out0["i_ids"], want["i_ids"]) and torch.equal(out0["j_ids"], want["j_ids"])
    # status codes on a live GPU
    with pytest.raises(_lib.PopeError):
        ops.coarse_match(torch.zeros(1, 16, 60, device=DEV), torch.zeros(1, 16, 60, device=DEV), (4, 4), (4, 4), 8.0)
    with pytest.raises(_lib.PopeError):
        ops.coarse_match(torch.zeros(1, 16, 64, device=DEV), torch.zeros(1, 16, 64, device=DEV), (4, 5), (4, 4), 8.0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("channels_last", [True, False])
def test_fine_gather_golden(golden_dir, dtype, channels_last):
    g = _load(golden_dir, "fine_preprocess")
    meta = json.loads(str(g["meta"]))
    case = COARSE_CASES[meta["coarse_case"]]
    c = _load(golden_dir, meta["coarse_case"])
    b_ids, i_ids, j_ids = (torch.from_numpy(c[k]).to(DEV) for k in ("b_ids", "i_ids", "j_ids"))
    hw = case["hw0_c"]
    ff0, ff1 = synth.fine_feature_maps(meta["fine_seed"], case["n"], hw[0] * 4, hw[1] * 4, 128, channels_last=channels_last)
    a, b = ff0.to(DEV, dtype), ff1.to(DEV, dtype)
    if channels_last:
        a, b = a.contiguous(memory_format=torch.channels_last), b.contiguous(memory_format=torch.channels_last)
    w0, w1 = ops.fine_gather(a, b, b_ids, i_ids, j_ids, hw[1], hw[1], 4, 5)
    assert w0.shape == (b_ids.numel(), 25, 128) and w0.dtype == dtype
    if dtype == torch.float32:       # a gather is bit-exact
        np.testing.assert_array_equal(w0[: meta["kept"]].cpu().numpy(), g["win0"])
        np.testing.assert_array_equal(w1[: meta["kept"]].cpu().numpy(), g["win1"])
        np.testing.assert_allclose(w0.sum((1, 2)).cpu().numpy(), g["win0_sum"], rtol=1e-5, atol=1e-4)
    else:
        want0 = torch.from_numpy(g["win0"]).to(torch.bfloat16)
        assert torch.equal(w0[: meta["kept"]].cpu(), want0)
        assert torch.equal(w1[: meta["kept"]].cpu(), torch.from_numpy(g["win1"]).to(torch.bfloat16))


def test_fine_gather_ragged_maps_and_device_count():
    """Different map sizes for the two images, windows hanging over every border, and the device-side count."""
    h0, w0, h1, w1, n = 7, 9, 5, 6, 3
    ff0, _ = synth.fine_feature_maps(3, n, h0 * 4, w0 * 4, 128)
    ff1, _ = synth.fine_feature_maps(4, n, h1 * 4, w1 * 4, 128)
    g = torch.Generator().manual_seed(9)
    M = 200
    b = torch.randint(0, n, (M,), generator=g).sort()[0]
    i = torch.randint(0, h0 * w0, (M,), generator=g)
    j = torch.randint(0, h1 * w1, (M,), generator=g)
    i[:4] = torch.tensor([0, w0 - 1, (h0 - 1) * w0, h0 * w0 - 1])
    want0, want1 = O.fine_windows(ff0, b, i), O.fine_windows(ff1, b, j)
    live = torch.tensor([150], dtype=torch.int32, device=DEV)
    w0_, w1_ = ops.fine_gather(ff0.to(DEV), ff1.to(DEV), b.to(DEV), i.to(DEV), j.to(DEV), w0, w1, 4, 5, m_dev=live)
    assert torch.equal(w0_[:150].cpu(), want0[:150]) and torch.equal(w1_[:150].cpu(), want1[:150])
    full0, full1 = ops.fine_gather(ff0.to(DEV), ff1.to(DEV), b.to(DEV), i.to(DEV), j.to(DEV), w0, w1, 4, 5)
    assert torch.equal(full0.cpu(), want0) and torch.equal(full1.cpu(), want1)


@pytest.mark.parametrize("name", ["fine_match_soft", "fine_match_peaked"])
def test_fine_match_golden(golden_dir, name):
    g = _load(golden_dir, name)
    meta = json.loads(str(g["meta"]))
    a, b = synth.fine_windows(meta["seed"], meta["M"], 25, 128, gain=meta["gain"])
    expec, mk1f = ops.fine_match(a.to(DEV), b.to(DEV), torch.from_numpy(g["mkpts1_c"]).to(DEV), 4.0)
    np.testing.assert_allclose(expec[:, :2].cpu().numpy(), g["expec_f"][:, :2], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(expec[:, 2].cpu().numpy(), g["expec_f"][:, 2], rtol=1e-4, atol=2e-3)   # cancellation
    np.testing.assert_allclose(mk1f.cpu().numpy(), g["mkpts1_f"], rtol=1e-4, atol=1e-5)
    # bf16 windows against the oracle on the rounded values
    a16, b16 = a.to(torch.bfloat16), b.to(torch.bfloat16)
    want = O.fine_match(a16.float(), b16.float(), torch.from_numpy(g["mkpts0_c"]), torch.from_numpy(g["mkpts1_c"]), 2.0)
    expec, mk1f = ops.fine_match(a16.to(DEV), b16.to(DEV), torch.from_numpy(g["mkpts1_c"]).to(DEV), 4.0)
    assert torch.allclose(expec[:, :2].cpu(), want["expec_f"][:, :2], rtol=1e-2, atol=1e-5)
    assert torch.allclose(mk1f.cpu(), want["mkpts1_f"], rtol=1e-2, atol=1e-4)


def test_fine_match_known_answers():
    M, WW, C = 25, 25, 128
    w0 = torch.zeros(M, WW, C); w1 = torch.zeros(M, WW, C)
    w0[:, 12, 0] = 100.0
    for r in range(25):
        w1[r, r, 0] = 100.0
    expec, mk1f = ops.fine_match(w0.to(DEV), w1.to(DEV), torch.zeros(M, 2, device=DEV), 4.0)
    lin = torch.linspace(-1, 1, 5)
    want = torch.stack([lin.repeat(5), lin.repeat_interleave(5)], 1)
    assert torch.allclose(expec[:, :2].cpu(), want, atol=1e-6)
    assert torch.allclose(mk1f.cpu(), want * 4.0, atol=1e-5)
    expec, _ = ops.fine_match(torch.zeros(3, WW, C, device=DEV), torch.zeros(3, WW, C, device=DEV),
                              torch.zeros(3, 2, device=DEV), 4.0)
    assert torch.allclose(expec[:, :2].cpu(), torch.zeros(3, 2), atol=1e-7)
    assert torch.allclose(expec[:, 2].cpu(), torch.full((3,), 2 * 0.5 ** 0.5), atol=1e-6)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_cosine_topk_vs_oracle(dtype):
    q, refs = synth.retrieval_tokens(7, 256, 384, dtype=dtype)
    want = O.cosine_scores(q, refs)
    ws, wi = O.running_topk(want.tolist(), 3)
    scores, slot_s, slot_i = ops.cosine_topk(q.to(DEV), refs.to(DEV), 3)
    assert torch.allclose(scores.cpu(), want, rtol=1e-5, atol=1e-6)
    assert slot_i.cpu().tolist() == wi
    assert torch.allclose(slot_s.cpu(), torch.tensor(ws), rtol=1e-5, atol=1e-6)
    # nothing positive -> all slots empty, like the loop that starts from zeros
    _, s, i = ops.cosine_topk(q.to(DEV), (-q).repeat(5, 1).to(DEV), 3)
    assert i.cpu().tolist() == [-1, -1, -1] and s.cpu().tolist() == [0.0, 0.0, 0.0]
    # the sharded form: crops scored in two blocks (as two ranks would), scores concatenated in crop order, top-k over all
    from pope_b200 import retrieval
    halves = torch.cat([ops.cosine_topk(q.to(DEV), refs[:100].to(DEV), 3)[0], ops.cosine_topk(q.to(DEV), refs[100:].to(DEV), 3)[0]])
    sc2, s2, i2 = retrieval.retrieve_topk_sharded(halves, 256, 0, 1, 3)
    assert i2 == wi and torch.equal(sc2, scores) and s2 == slot_s.cpu().tolist()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_hot_path_end_to_end_vs_oracle(dtype):
    h0, w0, h1, w1, n = 30, 40, 24, 32, 3
    f0, f1 = synth.coarse_features(51, n, h0 * w0, h1 * w1, 256, sigma=0.9, dtype=dtype)
    ff0, _ = synth.fine_feature_maps(52, n, h0 * 4, w0 * 4, 128, dtype=dtype)
    ff1, _ = synth.fine_feature_maps(53, n, h1 * 4, w1 * 4, 128, dtype=dtype)
    want = O.match_pairs(f0.float(), f1.float(), ff0.float(), ff1.float(), (h0 * 8, w0 * 8), (h0, w0), (h1, w1))
    for fused in (True, False):
        res = ops.match_pairs_device(f0.to(DEV), f1.to(DEV), ff0.to(DEV), ff1.to(DEV), (h0 * 8, w0 * 8), (h0, w0),
                                     (h1, w1), fused_fine=fused)
        m = res.total()
        assert m == want["b_ids"].numel()
        for k in ("b_ids", "i_ids", "j_ids"):
            assert torch.equal(res[k][:m].cpu(), want[k])
        tol = 1e-2 if dtype == torch.bfloat16 else 1e-4
        assert torch.allclose(res["mconf"][:m].cpu(), want["mconf"], rtol=tol, atol=0)
        assert torch.equal(res["mkpts0_f"][:m].cpu(), want["mkpts0_f"])
        assert torch.allclose(res["mkpts1_f"][:m].cpu(), want["mkpts1_f"], rtol=tol, atol=1e-3)
        assert torch.allclose(res["expec_f"][:m, :2].cpu(), want["expec_f"][:, :2], rtol=tol, atol=1e-4)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_fused_fine_equals_gather_then_match(dtype):
    """pope_fine_match_maps == pope_fine_gather + pope_fine_match, including border windows (bit for bit for fp32;
    for bf16 the fused kernel sums 8 channels per lane instead of 4, so only to fp32 rounding)."""
    h0, w0, h1, w1, n, M = 9, 11, 7, 8, 3, 500
    ff0, _ = synth.fine_feature_maps(81, n, h0 * 4, w0 * 4, 128, dtype=dtype)
    ff1, _ = synth.fine_feature_maps(82, n, h1 * 4, w1 * 4, 128, dtype=dtype)
    g = torch.Generator().manual_seed(83)
    b = torch.randint(0, n, (M,), generator=g).sort()[0].to(DEV)
    i = torch.randint(0, h0 * w0, (M,), generator=g).to(DEV)
    j = torch.randint(0, h1 * w1, (M,), generator=g).to(DEV)
    j[:4] = torch.tensor([0, w1 - 1, (h1 - 1) * w1, h1 * w1 - 1])
    mk1 = (torch.randint(0, 64, (M, 2), generator=g) * 8).float().to(DEV)
    a, c = ff0.to(DEV), ff1.to(DEV)
    w0_, w1_ = ops.fine_gather(a, c, b, i, j, w0, w1, 4, 5)
    e_ref, k_ref = ops.fine_match(w0_, w1_, mk1, 4.0)
    e_fused, k_fused = ops.fine_match_maps(a, c, b, i, j, mk1, w0, w1, 4, 4.0)
    if dtype == torch.float32:
        assert torch.equal(e_ref, e_fused) and torch.equal(k_ref, k_fused)
    else:
        assert torch.allclose(e_ref, e_fused, rtol=1e-5, atol=2e-6) and torch.allclose(k_ref, k_fused, rtol=1e-6, atol=1e-5)
    live = torch.tensor([123], dtype=torch.int32, device=DEV)
    e_part, _ = ops.fine_match_maps(a, c, b, i, j, mk1, w0, w1, 4, 4.0, m_dev=live)
    assert torch.equal(e_part[:123], e_fused[:123])
    # processing order sorted by reference cell: a permutation within each pair, results unchanged
    counts = torch.cat([torch.bincount(b, minlength=n).to(torch.int32), torch.tensor([M, 0], dtype=torch.int32, device=DEV)])
    order = ops.match_order_by_ref(counts, n, h1 * w1, j)
    assert sorted(order.tolist()) == list(range(M))
    key = (b[order.long()] * (h1 * w1) + j[order.long()]).cpu()
    assert bool((key[1:] >= key[:-1]).all())
    e_ord, k_ord = ops.fine_match_maps(a, c, b, i, j, mk1, w0, w1, 4, 4.0, order=order)
    assert torch.equal(e_ord, e_fused) and torch.equal(k_ord, k_fused)
    with pytest.raises(_lib.PopeError):     # plain NCHW maps are not accepted by the fused kernel
        ops.fine_match_maps(a.contiguous(), c.contiguous(), b, i, j, mk1, w0, w1, 4, 4.0)


def test_matcher_module_flow_matches_oracle():
    """The drop-in modules chained as Matcher.forward chains them (steps 3-5), weights shared with a CPU copy."""
    import copy
    import pope_b200
    torch.manual_seed(0)
    m_gpu = pope_b200.Matcher(pope_b200.make_default_cfg()).eval()
    m_cpu = copy.deepcopy(m_gpu)
    m_gpu = m_gpu.to(DEV)
    h, w, n = 20, 24, 2
    f0, f1 = synth.coarse_features(61, n, h * w, h * w, 256, sigma=0.85)
    ff0, ff1 = synth.fine_feature_maps(62, n, h * 4, w * 4, 128, channels_last=False)
    data = {"hw0_i": torch.Size([h * 8, w * 8]), "hw1_i": torch.Size([h * 8, w * 8]), "hw0_c": torch.Size([h, w]),
            "hw1_c": torch.Size([h, w]), "hw0_f": torch.Size([h * 4, w * 4]), "hw1_f": torch.Size([h * 4, w * 4]), "bs": n}
    with torch.no_grad():
        m_gpu.coarse_matching(f0.to(DEV), f1.to(DEV), data)
        w0, w1 = m_gpu.fine_preprocess(ff0.to(DEV), ff1.to(DEV), f0.to(DEV), f1.to(DEV), data)
        w0, w1 = m_gpu.loftr_fine(w0, w1)
        m_gpu.fine_matching(w0, w1, data)
        # CPU: oracle for the hot-path stages, the same torch modules for the Linears / transformer
        want = O.coarse_match(f0, f1, data["hw0_i"], (h, w), (h, w))
        c0, c1 = O.fine_windows(ff0, want["b_ids"], want["i_ids"]), O.fine_windows(ff1, want["b_ids"], want["j_ids"])
        cw = m_cpu.fine_preprocess.down_proj(torch.cat([f0[want["b_ids"], want["i_ids"]], f1[want["b_ids"], want["j_ids"]]], 0))
        mg = m_cpu.fine_preprocess.merge_feat(torch.cat([torch.cat([c0, c1], 0), cw[:, None].expand(-1, 25, -1)], -1))
        c0, c1 = m_cpu.loftr_fine(*mg.chunk(2, 0))
        wf = O.fine_match(c0, c1, want["mkpts0_c"], want["mkpts1_c"], 2.0)
    keys = list(data.keys())
    assert keys[7:] == ["b_ids", "i_ids", "j_ids", "gt_mask", "m_bids", "mkpts0_c", "mkpts1_c", "mconf", "W", "expec_f",
                        "mkpts0_f", "mkpts1_f"]
    for k in ("b_ids", "i_ids", "j_ids"):
        assert torch.equal(data[k].cpu(), want[k])
    assert data["W"] == 5
    assert torch.allclose(data["mconf"].cpu(), want["mconf"], rtol=1e-4)
    assert torch.allclose(data["mkpts1_f"].cpu(), wf["mkpts1_f"], rtol=1e-4, atol=2e-3)
    assert torch.equal(data["mkpts0_f"].cpu(), wf["mkpts0_f"])


def test_matcher_forward_images_and_empty_result():
    import pope_b200
    m = pope_b200.Matcher(pope_b200.make_default_cfg()).eval().to(DEV)
    batch = {"image0": torch.rand(2, 1, 96, 128, device=DEV), "image1": torch.rand(2, 1, 64, 64, device=DEV)}
    with torch.no_grad():
        assert m(batch) is None
    assert batch["mkpts0_f"].shape[1] == 2 and batch["mkpts0_f"].shape == batch["mkpts1_f"].shape
    assert batch["expec_f"].shape[1] == 3 and batch["b_ids"].dtype == torch.int64
    with torch.no_grad():
        fc0, fc1 = m({"image0": batch["image0"], "image1": batch["image1"]}, only_att_fea=True)
    assert fc0.shape == (2, 12 * 16, 256) and fc1.shape == (2, 8 * 8, 256)


def test_host_pipeline_equals_device_path():
    """pope_match_pairs_host (pinned host buffers, chunked streams) returns what the device path returns."""
    from pope_b200 import driver
    h, w, n = 20, 24, 5
    f0, f1 = synth.coarse_features(71, n, h * w, h * w, 256, sigma=0.85, dtype=torch.bfloat16)
    ff0, ff1 = synth.fine_feature_maps(72, n, h * 4, w * 4, 128, dtype=torch.bfloat16)
    res = ops.match_pairs_device(f0.to(DEV), f1.to(DEV), ff0.to(DEV), ff1.to(DEV), (h * 8, w * 8), (h, w), (h, w))
    m = res.total()
    # pageable host buffers (bulk copies of everything) ...
    out = driver.match_pairs_host(f0, f1, ff0, ff1, (h * 8, w * 8), (h, w), (h, w), chunk_pairs=2, device=0)
    # ... and page-locked ones: the fine kernel reads image 0's centre pixels in place over the host link; image 1's map is
    # either read window by window in place ("windows"), fetched once per pixel of the windows' union into the device map
    # ("union"), or copied whole ("bulk", also POPE_PIPELINE_WINDOWS_IN_PLACE=0)
    nhwc0, nhwc1 = ff0.permute(0, 2, 3, 1).contiguous().pin_memory(), ff1.permute(0, 2, 3, 1).contiguous().pin_memory()
    full = sum(t.numel() * t.element_size() for t in (f0, f1, nhwc0, nhwc1))
    need = np.zeros((n, h * 4, w * 4), dtype=bool)             # pixels of image 1 inside the window of some matched cell
    for b, j in zip(res["b_ids"][:m].tolist(), res["j_ids"][:m].tolist()):
        cy, cx = divmod(j, w)
        need[b, max(0, 4 * cy - 2):4 * cy + 3, max(0, 4 * cx - 2):4 * cx + 3] = True
    assert 0 < need.sum() < m * 25
    outs = [out]
    for var, mode in (("POPE_PIPELINE_F1", "windows"), ("POPE_PIPELINE_F1", "bulk"), ("POPE_PIPELINE_F1", "union"),
                      ("POPE_PIPELINE_WINDOWS_IN_PLACE", "0"), (None, None)):
        if var:
            os.environ[var] = mode
        try:
            pl = driver.Pipeline(torch.bfloat16, 2, (h * 8, w * 8), (h, w), (h, w), device=0)
            outs.append(pl.run(f0.pin_memory(), f1.pin_memory(), nhwc0, nhwc1))
            want = full - nhwc0.numel() * 2 + m * 128 * 2
            if var is None:
                assert pl.last_f1_mode in ("windows", "union")          # the default is one of the two in-place forms
            else:
                assert pl.last_f1_mode == ("bulk" if mode == "0" else mode)
            if pl.last_f1_mode == "windows":
                want += -nhwc1.numel() * 2 + m * 25 * 128 * 2
            elif pl.last_f1_mode == "union":
                want += -nhwc1.numel() * 2 + int(need.sum()) * 128 * 2
            assert pl.last_h2d_bytes == want, (mode, pl.last_h2d_bytes, want)
            pl.close()
        finally:
            if var:
                del os.environ[var]
    # fp32 maps (512-byte pixels: the other instantiation of the union fetch) against the bulk copy
    g0, g1 = synth.coarse_features(73, 3, h * w, h * w, 256, sigma=0.85, dtype=torch.float32)
    gg0, gg1 = synth.fine_feature_maps(74, 3, h * 4, w * 4, 128, dtype=torch.float32)
    q0, q1 = gg0.permute(0, 2, 3, 1).contiguous().pin_memory(), gg1.permute(0, 2, 3, 1).contiguous().pin_memory()
    got = {}
    for mode in ("bulk", "union"):
        os.environ["POPE_PIPELINE_F1"] = mode
        try:
            pl = driver.Pipeline(torch.float32, 2, (h * 8, w * 8), (h, w), (h, w), device=0)
            got[mode] = pl.run(g0.pin_memory(), g1.pin_memory(), q0, q1)
            assert pl.last_f1_mode == mode
            pl.close()
        finally:
            del os.environ["POPE_PIPELINE_F1"]
    assert int(got["bulk"]["counts"].sum()) > 0
    assert torch.equal(got["bulk"]["counts"], got["union"]["counts"])
    fa, fb = driver.flatten_slots(got["bulk"]), driver.flatten_slots(got["union"])
    for k in ("i_ids", "j_ids", "mconf", "mkpts0_f", "mkpts1_f"):
        assert torch.equal(fa[k], fb[k]), k
    for o in outs:
        assert int(o["counts"].sum()) == m
        cat = driver.flatten_slots(o)
        for k in ("b_ids", "i_ids", "j_ids"):
            assert torch.equal(cat[k], res[k][:m].cpu()), k
        assert torch.equal(cat["mconf"], res["mconf"][:m].cpu())
        assert torch.equal(cat["mkpts0_f"], res["mkpts0_f"][:m].cpu())
        assert torch.equal(cat["mkpts1_f"], res["mkpts1_f"][:m].cpu())


def test_match_scores_vs_oracle():
    """Per-pair count of mconf > 0.9 and the per-query arg-max (eval_linemod_json.py:118-119,146) on the device list."""
    n, h, w = 7, 20, 24
    f0, f1 = synth.coarse_features(71, n, h * w, h * w, 256, sigma=0.95)
    res = ops.coarse_match(f0.to(DEV), f1.to(DEV), (h, w), (h, w), 8.0)
    scores, best = ops.match_scores(res["mconf"], res["counts"], n, group=3, thr=0.9)
    out = res.sliced()
    per_pair = [out["mconf"][out["b_ids"] == p].cpu().numpy() for p in range(n)]
    want_s, want_b = O.match_scores(per_pair, group=3, thr=0.9)
    assert scores.tolist() == want_s and best.tolist() == want_b
    assert sum(want_s) > 0
    # ties: equal scores -> the first crop wins, like np.argmax
    conf = torch.tensor([0.95, 0.95, 0.5, 0.95, 0.95, 0.1], device=DEV)
    cnt = torch.tensor([2, 1, 2, 1, 6, 0], dtype=torch.int32, device=DEV)      # pairs 0..3 (+ total, flags)
    s2, b2 = ops.match_scores(conf, cnt, 4, group=4, thr=0.9)
    assert s2.tolist() == [2, 0, 2, 0] and b2.tolist() == [0]
    # no match in any pair: an empty list (no storage behind it) is valid input
    s3, b3 = ops.match_scores(torch.zeros(0, device=DEV), torch.zeros(5, dtype=torch.int32, device=DEV), 5, group=3, thr=0.9)
    assert s3.tolist() == [0] * 5 and b3.tolist() == [0, 0]


def test_pack_records_kernel_equals_torch_packing():
    """pope_pack_records (one kernel, live count read on the device) against the torch cat it replaced."""
    from pope_b200 import driver
    n, h, w = 3, 20, 24
    f0, f1 = synth.coarse_features(72, n, h * w, h * w, 256, sigma=0.9)
    ff0, ff1 = synth.fine_feature_maps(73, n, h * 4, w * 4, 128, channels_last=True)
    res = ops.match_pairs_device(f0.to(DEV), f1.to(DEV), ff0.to(DEV), ff1.to(DEV), (h * 8, w * 8), (h, w), (h, w))
    m = res.total()
    rec = driver.pack_records(res, 40)
    want = torch.cat([(res["b_ids"] + 40).to(torch.int32)[:, None], res["i_ids"].to(torch.int32)[:, None],
                      res["j_ids"].to(torch.int32)[:, None], res["mconf"].view(torch.int32)[:, None],
                      res["mkpts0_f"].view(torch.int32), res["mkpts1_f"].view(torch.int32)], 1)
    assert m > 100 and rec.shape == want.shape and torch.equal(rec[:m], want[:m])
    # the compact 20-byte form (the keypoint of image 0 is implied by i) decodes to the same lists
    full = driver.unpack_records(rec[:m].cpu())
    comp = driver.unpack_records(driver.pack_records(res, 40, compact=True)[:m].cpu(), w, 8.0)
    for k in full:
        assert torch.equal(full[k], comp[k]), k
    assert torch.equal(comp["mkpts0_f"], res["mkpts0_f"][:m].cpu()) and torch.equal(comp["i_ids"], res["i_ids"][:m].cpu())


def test_job_gather_appends_from_alternating_streams():
    """JobGather.add orders appends issued from different streams by itself (the usage DeviceBatchRunner.submit(after=)
    invites): four batches on two alternating streams end up behind one another, none overwritten, count exact."""
    from pope_b200 import driver
    n, h, w = 2, 20, 24
    ff0, ff1 = synth.fine_feature_maps(75, n, h * 4, w * 4, 128, channels_last=True)
    ff0, ff1 = ff0.to(DEV), ff1.to(DEV)
    runner = driver.DeviceBatchRunner(DEV, 2)
    job = driver.JobGather(4, n * h * w, DEV, compact=True)
    feats = [tuple(t.to(DEV) for t in synth.coarse_features(76 + k, n, h * w, h * w, 256, sigma=0.9)) for k in range(4)]
    runner.fork()
    results = [runner.submit(f0, f1, ff0, ff1, (h * 8, w * 8), (h, w), (h, w), after=lambda r, k=k: job.add(r, 10 * k))[0]
               for k, (f0, f1) in enumerate(feats)]
    runner.join()
    recs, sizes = job.finish()
    torch.cuda.synchronize()
    ms = [r.total() for r in results]
    assert sizes == [sum(ms)] and min(ms) > 50
    got = driver.unpack_records(recs[0].cpu(), w, 8.0)
    off = 0
    for k, r in enumerate(results):
        m = ms[k]
        assert torch.equal(got["b_ids"][off:off + m], r["b_ids"][:m].cpu() + 10 * k)
        assert torch.equal(got["i_ids"][off:off + m], r["i_ids"][:m].cpu())
        assert torch.equal(got["mkpts1_f"][off:off + m], r["mkpts1_f"][:m].cpu())
        off += m


def test_coarse_many_pairs_two_launch_compaction():
    """More pairs than co-resident compaction CTAs (n > 2 x SMs): count and emit run as two launches; same lists as the
    pairs processed in small batches."""
    n, h, w = 320, 8, 10
    f0, f1 = synth.coarse_features(81, n, h * w, h * w, 64, sigma=0.8, dtype=torch.bfloat16)
    big = _run_coarse(f0.float(), f1.float(), (h, w), (h, w), IMPLS["tcgen05"], torch.bfloat16)
    parts = [_run_coarse(f0[a:a + 64].float(), f1[a:a + 64].float(), (h, w), (h, w), IMPLS["tcgen05"], torch.bfloat16)
             for a in range(0, n, 64)]
    assert big["b_ids"].numel() > 1000
    assert torch.equal(big["b_ids"], torch.cat([p["b_ids"] + 64 * k for k, p in enumerate(parts)]))
    for key in ("i_ids", "j_ids", "mconf"):
        assert torch.equal(big[key], torch.cat([p[key] for p in parts])), key


def test_device_batch_runner_equals_sequential_steps():
    """driver.DeviceBatchRunner: batches on alternating streams give the results of one-at-a-time calls."""
    from pope_b200 import driver
    h, w = 20, 24
    batches = []
    for seed in (91, 92, 93):
        f0, f1 = synth.coarse_features(seed, 2, h * w, h * w, 256, sigma=0.9, dtype=torch.bfloat16)
        ff0, ff1 = synth.fine_feature_maps(seed + 10, 2, h * 4, w * 4, 128, dtype=torch.bfloat16)
        batches.append([t.to(DEV) for t in (f0, f1, ff0, ff1)])
    runner = driver.DeviceBatchRunner(DEV, 2)
    runner.fork()
    got = [runner.submit(*b, (h * 8, w * 8), (h, w), (h, w))[0] for b in batches]
    runner.join()
    torch.cuda.synchronize()
    for b, res in zip(batches, got):
        want = ops.match_pairs_device(*b, (h * 8, w * 8), (h, w), (h, w))
        m = want.total()
        assert res.total() == m and m > 100
        for k in ("b_ids", "i_ids", "j_ids", "mconf", "mkpts1_f", "expec_f"):
            assert torch.equal(res[k][:m], want[k][:m]), k


def test_hot_path_step_is_graph_capturable():
    """ops.match_pairs_device has no host synchronisation, so a step can be captured into a CUDA graph and replayed on new
    feature contents.  24 pairs at 40x48 cells take the head / tail sweep with the programmatically serialised merge launch
    (csrc/coarse_tc.cu::coarse_tc_run), which must survive capture; the replays equal the eager calls bit for bit."""
    n, h, w = 24, 40, 48
    assert _lib.single_sweep_is_split(n, h * w)
    _need_tc("tcgen05", 256, h * w, h * w)
    sets = []
    for seed in (201, 202):
        f0, f1 = synth.coarse_features(seed, n, h * w, h * w, 256, sigma=1.0, dtype=torch.bfloat16)
        ff0, ff1 = synth.fine_feature_maps(seed + 10, n, h * 4, w * 4, 128, dtype=torch.bfloat16, channels_last=True)
        sets.append([t.to(DEV) for t in (f0, f1, ff0, ff1)])
    keys = ("b_ids", "i_ids", "j_ids", "mconf", "mkpts1_f", "expec_f")
    want = []
    for s in sets:
        r = ops.match_pairs_device(*s, (h * 8, w * 8), (h, w), (h, w))
        m = r.total()
        assert m > 5000 and r.flags() == 0
        want.append((m, {k: r[k][:m].clone() for k in keys}))
    static = [t.clone() for t in sets[0]]
    ws = torch.empty(_lib.lib().pope_coarse_workspace_bytes_ex(n, h * w, h * w, 256, 1), dtype=torch.uint8, device=DEV)
    side = torch.cuda.Stream(DEV)
    side.wait_stream(torch.cuda.current_stream(DEV))
    with torch.cuda.stream(side):                      # warm-up on the capture stream (first-use attribute calls)
        ops.match_pairs_device(*static, (h * 8, w * 8), (h, w), (h, w), workspace=ws)
    torch.cuda.current_stream(DEV).wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        res = ops.match_pairs_device(*static, (h * 8, w * 8), (h, w), (h, w), workspace=ws)
    for which in (1, 0, 1):
        for dst, src in zip(static, sets[which]):
            dst.copy_(src)
        graph.replay()
        torch.cuda.synchronize()
        m, ref = want[which]
        assert res.total() == m and res.flags() == 0
        for k in keys:
            assert torch.equal(res[k][:m], ref[k]), (which, k)


def test_match_crops_equals_the_reference_pair_loop():
    """driver.match_crops against the loop of eval_linemod_json.py:103-122 / :146 restated with cv2 + numpy.  A stand-in
    matcher whose output is a deterministic function of its two input images makes the comparison non-vacuous (the real
    Matcher with random weights finds no match); the real Matcher then runs through the same helper for the flow."""
    cv2 = pytest.importorskip("cv2")
    import pope_b200
    from pope_b200 import driver

    def fake_matcher(batch):
        n = batch["image0"].shape[0]
        b, conf, k0, k1 = [], [], [], []
        for p in range(n):
            key1 = int((batch["image1"][p] * 255).round().long().sum().item())       # exact, whatever the batching
            key0 = int((batch["image0"][p] * 255).round().long().sum().item())
            m = key1 % 40 + 3
            s = float(key0 % 7)
            t = torch.arange(m, device=batch["image1"].device, dtype=torch.float32)
            b.append(torch.full((m,), p, device=t.device, dtype=torch.int64))
            conf.append(((t * 37 + key1 % 100) % 100) / 100.0)
            k0.append(torch.stack([t * s, t + 1], 1))
            k1.append(torch.stack([t + 2, t * 3], 1))
        batch.update(b_ids=torch.cat(b), mconf=torch.cat(conf), mkpts0_f=torch.cat(k0), mkpts1_f=torch.cat(k1))

    rng = np.random.default_rng(8)
    image0 = rng.integers(0, 256, (48, 64, 3), dtype=np.uint8)
    crops = [rng.integers(0, 256, s + (3,), dtype=np.uint8) for s in ((32, 40), (24, 24), (32, 40), (32, 40), (24, 24))]
    res, scores, best = driver.match_crops(fake_matcher, torch.from_numpy(image0).to(DEV), [torch.from_numpy(c).to(DEV) for c in crops])
    want_scores = []
    for k, c in enumerate(crops):                              # the reference loop, one call per crop
        g0 = torch.from_numpy(cv2.cvtColor(image0, cv2.COLOR_BGR2GRAY)).float()[None] / 255.
        g1 = torch.from_numpy(cv2.cvtColor(c, cv2.COLOR_BGR2GRAY)).float()[None] / 255.
        batch = {"image0": g0.unsqueeze(0).to(DEV), "image1": g1.unsqueeze(0).to(DEV)}
        fake_matcher(batch)
        conf = batch["mconf"].cpu().numpy()
        want_scores.append(int(np.where(conf > 0.9)[0].shape[0]))
        assert torch.equal(res[k]["mconf"], batch["mconf"]) and torch.equal(res[k]["mkpts0_f"], batch["mkpts0_f"])
        assert torch.equal(res[k]["mkpts1_f"], batch["mkpts1_f"])
    assert scores.tolist() == want_scores and best == int(np.argmax(want_scores)) and max(want_scores) > 0
    # flow with the real module (random weights: no matches, every score 0, first crop wins like np.argmax)
    torch.manual_seed(0)
    matcher = pope_b200.Matcher(pope_b200.make_default_cfg()).eval().to(DEV)
    res, scores, best = driver.match_crops(matcher, torch.from_numpy(image0).to(DEV), [torch.from_numpy(c).to(DEV) for c in crops[:3]])
    assert len(res) == 3 and scores.tolist() == [0, 0, 0] and best == 0 and res[1]["mkpts0_f"].shape == (0, 2)


@pytest.mark.parametrize("impl", list(IMPLS))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_kat_border_veto_and_threshold_straddle(impl, dtype):
    """Known answers of SURVEY.md 8(c) on the CUDA path: (2) a border cell that holds the row maximum vetoes the interior
    candidate of its row and is itself removed; (3) two equal-strength columns split the row softmax, which puts the
    confidence on either side of the threshold depending on the third column's weight."""
    _need_tc(impl, 64, 64, 64)
    h = w = 8
    L = h * w
    i, j_in, j_bd = 3 * w + 3, 4 * w + 4, 0 * w + 5
    f0, f1 = torch.zeros(1, L, 64), torch.zeros(1, L, 64)
    f0[0, i, 0] = 16.0
    f1[0, j_in, 0], f1[0, j_bd, 0] = 8.0, 8.1875                              # exactly representable in bf16
    conf = O.dual_softmax_conf(f0, f1)
    assert conf[0, i, j_bd] > conf[0, i, j_in] > 0.2                          # the interior candidate alone would pass thr
    out = _run_coarse(f0, f1, (h, w), (h, w), IMPLS[impl], dtype, thr=0.2)
    assert out["i_ids"].numel() == 0                                          # vetoed by the border cell, which is removed
    out = _run_coarse(f0, f1, (h, w), (h, w), IMPLS[impl], dtype, thr=0.2, border_rm=0)
    assert out["i_ids"].tolist() == [i] and out["j_ids"].tolist() == [j_bd]   # without border removal the border cell matches
    f1[0, j_bd, 0] = 6.0                                                      # now the interior cell wins
    out = _run_coarse(f0, f1, (h, w), (h, w), IMPLS[impl], dtype, thr=0.2)
    want = O.coarse_match(f0, f1, (64, 64), (h, w), (h, w), thr=0.2)
    assert out["i_ids"].tolist() == [i] and out["j_ids"].tolist() == [j_in] == want["j_ids"].tolist()
    assert torch.allclose(out["mconf"], want["mconf"], rtol=1e-2 if dtype == torch.bfloat16 else 1e-4, atol=0)
    # (3) one query row, two interior columns of equal strength: no unique row maximum pair above thr unless one is lowered
    f0, f1 = torch.zeros(1, L, 64), torch.zeros(1, L, 64)
    ja, jb = 2 * w + 2, 5 * w + 5
    f0[0, i, 0] = 16.0
    for strength_b, n_expected in ((16.0, None), (8.0, 1)):
        f1[0, ja, 0], f1[0, jb, 0] = 16.0, strength_b
        want = O.coarse_match(f0, f1, (64, 64), (h, w), (h, w), thr=0.2)
        out = _run_coarse(f0, f1, (h, w), (h, w), IMPLS[impl], dtype, thr=0.2)
        if n_expected is None:      # an exact tie: both columns hold the row maximum; `mask.max(dim=2)` keeps the first
            assert out["i_ids"].tolist() == want["i_ids"].tolist() and out["j_ids"].tolist() == want["j_ids"].tolist()
        else:
            assert out["j_ids"].tolist() == [ja] == want["j_ids"].tolist()
            assert torch.allclose(out["mconf"], want["mconf"], rtol=1e-2 if dtype == torch.bfloat16 else 1e-4, atol=0)
