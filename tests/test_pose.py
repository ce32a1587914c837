"""Batched relative pose: oracle/pose_oracle.py against the fixture made with the unmodified
reference `estimate_pose` (cv2), the device math compiled for the host against the oracle bit for bit, and -- on a GPU --
pope_estimate_pose_batch against the oracle bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import pose_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "pose_cv2.npz")


def rot_angle(a, b):
    return float(np.degrees(np.arccos(np.clip((np.trace(a.T @ b) - 1.0) / 2.0, -1.0, 1.0))))


def dir_angle(a, b):
    return float(np.degrees(np.arccos(np.clip(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)), -1.0, 1.0))))


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


@pytest.fixture(scope="module")
def oracle_out(gold):
    return O.estimate_pose_batch(gold["mkpts0"], gold["mkpts1"], gold["counts"], gold["K0"], gold["K1"],
                                 float(gold["thresh"]), float(gold["conf_hi"]), 1000, seed=0)


PLANAR, SMALL_BASELINE = (10, 12), (11,)       # scenes appended by oracle/gen_golden_pose.py in round 2


@pytest.mark.parametrize("tag", ["hi", "lo"])
def test_oracle_against_reference_cv2(gold, oracle_out, tag):
    """The pin: same scenes through the reference's estimate_pose (metrics.py:69-94) at its default confidence 0.99999
    ("hi") and at the 0.99 of eval_onepose_json.py:164 ("lo").  RANSAC draws differ and neither side refines the winning
    minimal-sample model, so the bar is statistical: same None / not-None, rotations within 2 (3) degrees of cv2's and 1.5
    (2) of the truth, inlier masks agreeing with cv2's on more than 94 % (90 %) of the matches, precision >= 99.5 % and
    recall >= 93 % against the correspondences that were planted.  At 0.99 cv2 stops after 5-78 draws and sometimes keeps a
    model that misses a fifth of the planted matches (pairs 0 and 6); there the masks may differ by more, but only if the
    oracle recalls at least 5 points of a hundred more of the planted matches than cv2 does.
    Near-degenerate geometry: planar scenes have a two-fold ambiguity that costs both sides accuracy (cv2 is 6.7 degrees off
    on pair 10), so only the masks and a 6 degree bar against the truth are held; with a baseline of 1/225 of the depth
    (pair 11) recoverPose's cheirality vote is decided by noise -- cv2 returns one inlier / None -- so the oracle is held to
    the truth alone there (documented divergence)."""
    conf = float(gold[f"conf_{tag}"])
    out = oracle_out if tag == "hi" else O.estimate_pose_batch(gold["mkpts0"], gold["mkpts1"], gold["counts"], gold["K0"],
                                                              gold["K1"], float(gold["thresh"]), conf, 1000, seed=0)
    r_cv, r_gt, agree = (2.0, 1.5, 0.94) if tag == "hi" else (3.0, 2.0, 0.90)
    off = np.concatenate([[0], np.cumsum(gold["counts"])])
    for p, m in enumerate(gold["counts"]):
        mask_o, mask_c = out["inliers"][off[p]:off[p + 1]].astype(bool), gold[f"inliers_cv2_{tag}"][off[p]:off[p + 1]].astype(bool)
        planted = gold["planted"][off[p]:off[p + 1]]
        if p in SMALL_BASELINE:
            assert out["status"][p] == 1 and rot_angle(out["R"][p], gold["R_gt"][p]) < 1.0
            assert (mask_o & planted).sum() >= 0.995 * mask_o.sum() and (mask_o & planted).sum() >= 0.80 * planted.sum()
            continue
        assert out["status"][p] == gold[f"status_cv2_{tag}"][p]
        if not gold[f"status_cv2_{tag}"][p]:
            assert out["n_inliers"][p] == 0 and not mask_o.any()
            continue
        if m == 5:      # a bare minimal sample has several exact solutions; only the count is comparable
            assert mask_o.sum() == mask_c.sum() == 5
            continue
        if p in PLANAR:
            assert rot_angle(out["R"][p], gold["R_gt"][p]) < 6.0
        else:
            assert rot_angle(out["R"][p], gold[f"R_cv2_{tag}"][p]) < r_cv
            assert rot_angle(out["R"][p], gold["R_gt"][p]) < r_gt
            assert dir_angle(out["t"][p], gold["t_gt"][p]) < 6.0
        rec_o, rec_c = (mask_o & planted).sum() / planted.sum(), (mask_c & planted).sum() / planted.sum()
        assert (mask_o & planted).sum() >= 0.995 * mask_o.sum() and rec_o >= 0.93, (p, rec_o)
        assert np.mean(mask_o == mask_c) > agree or rec_o >= rec_c + 0.05, (p, np.mean(mask_o == mask_c), rec_o, rec_c)
        assert int(mask_o.sum()) >= 0.9 * int(mask_c.sum())
        assert abs(np.linalg.det(out["R"][p]) - 1.0) < 1e-9 and abs(np.linalg.norm(out["t"][p]) - 1.0) < 1e-9


def test_five_point_models_satisfy_constraints():
    rng = np.random.default_rng(5)
    x0, y0, x1, y1 = (rng.uniform(-0.5, 0.5, (32, 5)) for _ in range(4))
    models, n = O.five_point(x0, y0, x1, y1)
    assert n.max() <= 10 and n.sum() > 32
    for b in range(32):
        for r in range(n[b]):
            E = models[b, r].reshape(3, 3)
            a = np.stack([x1[b], y1[b], np.ones(5)], 1)
            c = np.stack([x0[b], y0[b], np.ones(5)], 1)
            assert np.abs(np.einsum("ki,ij,kj->k", a, E, c)).max() < 1e-7
            s = np.linalg.svd(E, compute_uv=False)
            assert abs(s[0] - s[1]) < 1e-6 and s[2] < 1e-6
        assert not models[b, n[b]:].any()


def test_five_point_finds_the_real_roots():
    """Against numpy's companion-matrix roots: the grid + bisection root finder loses at most a few close pairs."""
    rng = np.random.default_rng(6)
    c = rng.normal(size=(200, 11))
    roots, n = O.real_roots10(c)
    want = 0
    for b in range(200):
        r = np.roots(c[b, ::-1])
        real = np.sort(r[np.abs(r.imag) < 1e-9].real)
        want += len(real)
        got = np.sort(roots[b, :n[b]])
        assert len(got) <= len(real)
        for g in got:
            assert np.min(np.abs(real - g) / np.maximum(1.0, np.abs(real))) < 1e-8
    assert n.sum() >= 0.97 * want


def test_update_num_iters():
    assert O.update_num_iters(0.99, 0.0, 1000) == 0
    assert O.update_num_iters(0.99, 1.0, 1000) == 1000
    assert O.update_num_iters(0.99, 0.5, 1000) == int(np.rint(np.log(0.01) / np.log(1 - 0.5 ** 5)))
    assert O.update_num_iters(0.99999, 0.9, 1000) == 1000


def test_fewer_than_five_matches_is_none():
    k = np.eye(3)
    out = O.estimate_pose(np.zeros((4, 2), np.float32), np.zeros((4, 2), np.float32), k, k, 0.5)
    assert out["status"] == 0 and out["iters"] == 0


def test_oracle_chain_recovers_planted_pose():
    """The CPU restatement of the whole chain on geometry-consistent synthetic features (synth.posed_pair_features): the
    match oracle finds the planted correspondences within the 2 px fine grid and the pose oracle gets the motion back."""
    from oracle import pope_oracle as MO
    from pope_b200 import synth
    n, hw = 2, (30, 40)
    d = synth.posed_pair_features(3, n, hw_c=hw)
    w = MO.match_pairs(d["feat_c0"], d["feat_c1"], d["feat_f0"].float(), d["feat_f1"].float(), (240, 320), hw, hw)
    planted = int((~torch.isnan(d["proj"][..., 0])).sum())
    assert len(w["b_ids"]) > 0.6 * planted
    resid = (w["mkpts1_f"].double() - d["proj"][w["b_ids"], w["i_ids"]]).norm(dim=1)
    assert int(resid.isnan().sum()) == 0 and float(resid.median()) < 1.0
    counts = np.bincount(w["b_ids"].numpy(), minlength=n).astype(np.int32)
    K = np.tile(d["K"].numpy(), (n, 1, 1))
    o = O.estimate_pose_batch(w["mkpts0_f"].numpy(), w["mkpts1_f"].numpy(), counts, K, K, 1.0, 0.99999, 1000, 0)
    assert o["status"].tolist() == [1, 1]
    for p in range(n):
        assert rot_angle(o["R"][p], d["R"][p].numpy()) < 2.0
        assert o["n_inliers"][p] > 0.7 * counts[p]


@pytest.fixture(scope="module")
def host_math(tmp_path_factory):
    """pope_b200/csrc/pose_math.cuh compiled for the host (the file the device runs), without multiply-add contraction."""
    out = tmp_path_factory.mktemp("pm") / "libpm.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "pope_b200", "csrc"),
                    "-o", str(out), os.path.join(ROOT, "tests", "hostbuild", "pose_host.cpp")], check=True)
    L = C.CDLL(str(out))
    L.pm_five_point.argtypes = [C.c_void_p] * 5
    L.pm_decompose.argtypes = [C.c_void_p] * 4
    L.pm_cheirality.argtypes = [C.c_void_p, C.c_void_p] + [C.c_double] * 5
    L.pm_sampson.argtypes = [C.c_void_p] + [C.c_double] * 5
    return L


def test_device_math_source_equals_oracle_bit_for_bit(gold, oracle_out, host_math):
    L = host_math
    m = int(gold["counts"][1])
    off = int(gold["counts"][0])
    pts = O.normalize_points(gold["mkpts0"][off:off + m], gold["mkpts1"][off:off + m], gold["K0"][1], gold["K1"][1])
    hs = np.arange(96)
    idx, ok = O.draw5(11, 4, hs, m)
    for h in hs:
        got = (C.c_int * 5)()
        assert bool(L.pm_draw5(C.c_uint64(11), C.c_uint64(4), C.c_uint64(int(h)), m, got)) == bool(ok[h])
        assert list(got) == list(idx[h])
    models, n = O.five_point(pts[idx, 0], pts[idx, 1], pts[idx, 2], pts[idx, 3])
    assert n.sum() > 96
    for h in hs:
        cols = [np.ascontiguousarray(pts[idx[h], k]) for k in range(4)]
        got = np.zeros((10, 9))
        assert L.pm_five_point(*[c.ctypes.data for c in cols], got.ctypes.data) == n[h]
        assert np.array_equal(got[:n[h]], models[h, :n[h]])
    E = np.ascontiguousarray(oracle_out["E"][1].reshape(9))
    R1, R2, t = np.zeros(9), np.zeros(9), np.zeros(3)
    L.pm_decompose(E.ctypes.data, R1.ctypes.data, R2.ctypes.data, t.ctypes.data)
    w1, w2, wt = O.decompose_essential(E)
    assert np.array_equal(R1, w1) and np.array_equal(R2, w2) and np.array_equal(t, wt)
    thr2 = 1e-6
    inl = O.sampson_inlier(E, pts, thr2)
    assert 0 < inl.sum() < m
    for i in range(m):
        assert bool(L.pm_sampson(E.ctypes.data, *pts[i], thr2)) == bool(inl[i])
    for Rm, tv in ((R1, t), (R2, t), (R1, -t), (R2, -t)):
        tv = np.ascontiguousarray(tv)
        want = O.cheirality(Rm, tv, pts)
        got = np.array([L.pm_cheirality(Rm.ctypes.data, tv.ctypes.data, *pts[i], 1e9) for i in range(m)], dtype=bool)
        assert np.array_equal(want, got)


# ---- GPU: the C ABI against the oracle ------------------------------------------------------------------------------------

def _run_gpu(gold, seed=0, max_iters=1000, conf=None):
    from pope_b200 import pose
    dev = torch.device("cuda:0")
    return pose.estimate_pose_batch(torch.from_numpy(gold["mkpts0"]).to(dev), torch.from_numpy(gold["mkpts1"]).to(dev),
                                    torch.from_numpy(gold["counts"]).to(dev), torch.from_numpy(gold["K0"]),
                                    torch.from_numpy(gold["K1"]), float(gold["thresh"]),
                                    float(gold["conf_hi"]) if conf is None else conf, max_iters, seed)


@pytest.mark.gpu
def test_gpu_pose_equals_oracle_bit_for_bit(gold, oracle_out):
    got = _run_gpu(gold)
    for k in ("status", "iters", "n_inliers"):
        assert np.array_equal(got[k].cpu().numpy(), oracle_out[k]), k
    assert np.array_equal(got["inliers"].cpu().numpy(), oracle_out["inliers"])
    for k in ("E", "R", "t"):
        assert np.array_equal(got[k].cpu().numpy(), oracle_out[k]), (k, np.abs(got[k].cpu().numpy() - oracle_out[k]).max())


@pytest.mark.gpu
def test_gpu_pose_other_seed_and_bounds_against_oracle(gold):
    for seed, max_iters, conf in ((7, 40, 0.99999), (123456789, 1024, 0.999)):
        want = O.estimate_pose_batch(gold["mkpts0"], gold["mkpts1"], gold["counts"], gold["K0"], gold["K1"],
                                     float(gold["thresh"]), conf, max_iters, seed=seed)
        got = _run_gpu(gold, seed, max_iters, conf)
        for k in ("status", "iters", "n_inliers", "inliers", "E", "R", "t"):
            assert np.array_equal(got[k].cpu().numpy(), want[k]), (seed, k)


@pytest.mark.gpu
def test_gpu_pose_recovers_ground_truth_on_a_full_batch():
    """64 pairs x ~2 000 matches, 30 % outliers: every rotation within 2 degrees of the truth (median under 0.5; the winning
    minimal-sample model is not refined, as in OpenCV), deterministic in the seed."""
    from oracle.gen_golden_pose import scene
    from pope_b200 import pose
    rng = np.random.default_rng(3)
    sc = [scene(rng, int(rng.integers(1500, 2500)), 0.3, 0.1, 600.0, 650.0) for _ in range(64)]
    dev = torch.device("cuda:0")
    args = (torch.from_numpy(np.concatenate([s[0] for s in sc])).to(dev), torch.from_numpy(np.concatenate([s[1] for s in sc])).to(dev),
            torch.tensor([len(s[0]) for s in sc], dtype=torch.int32, device=dev), torch.from_numpy(np.stack([s[2] for s in sc])),
            torch.from_numpy(np.stack([s[3] for s in sc])), 0.5, 0.99999)
    a, b = pose.estimate_pose_batch(*args), pose.estimate_pose_batch(*args)
    for k in ("R", "t", "inliers", "iters"):
        assert torch.equal(a[k], b[k])
    assert bool((a["status"] == 1).all())
    R, t = a["R"].cpu().numpy(), a["t"].cpu().numpy()
    errs = [rot_angle(R[p], sc[p][4]) for p in range(64)]
    terrs = [dir_angle(t[p], sc[p][5]) for p in range(64)]
    assert max(errs) < 2.0 and np.median(errs) < 0.5 and np.median(terrs) < 2.0, (max(errs), np.median(errs), np.median(terrs))
    frac = a["n_inliers"].cpu().numpy() / np.array([len(s[0]) for s in sc])
    assert frac.min() > 0.5


@pytest.mark.gpu
def test_estimate_pose_drop_in_signature(gold):
    from pope_b200 import pose
    m0 = int(gold["counts"][0])
    ret = pose.estimate_pose(gold["mkpts0"][:m0], gold["mkpts1"][:m0], gold["K0"][0], gold["K1"][0], 0.5, 0.99)
    R, t, mask = ret
    assert R.shape == (3, 3) and t.shape == (3,) and mask.shape == (m0,) and mask.dtype == bool
    assert rot_angle(R, gold["R_gt"][0]) < 1.0
    assert pose.estimate_pose(gold["mkpts0"][:4], gold["mkpts1"][:4], gold["K0"][0], gold["K1"][0], 0.5) is None


@pytest.mark.gpu
def test_pose_consumes_the_hot_path_output_without_leaving_the_device():
    """match_pairs_device -> estimate_pose_batch on the capacity-sized device arrays (no host sync in between); the
    oracle gets the same match lists."""
    from pope_b200 import ops, pose, synth
    dev = torch.device("cuda:0")
    n, h, w = 3, 24, 32
    f0, f1 = synth.coarse_features(11, n, h * w, h * w, 256, sigma=0.9, dtype=torch.bfloat16)
    ff0, ff1 = synth.fine_feature_maps(12, n, h * 4, w * 4, 128, dtype=torch.bfloat16)
    res = ops.match_pairs_device(f0.to(dev), f1.to(dev), ff0.to(dev), ff1.to(dev), (h * 8, w * 8), (h, w), (h, w))
    K = torch.tensor([[300.0, 0, w * 4.0], [0, 300.0, h * 4.0], [0, 0, 1]], dtype=torch.float64).expand(n, 3, 3)
    got = pose.estimate_pose_batch(res["mkpts0_f"], res["mkpts1_f"], res["counts"], K, K, 0.5, 0.99, max_iters=64)
    counts = res["counts"][:n].cpu().numpy()
    m = int(counts.sum())
    assert m > 100
    want = O.estimate_pose_batch(res["mkpts0_f"][:m].cpu().numpy(), res["mkpts1_f"][:m].cpu().numpy(), counts, K.numpy(),
                                 K.numpy(), 0.5, 0.99, 64, seed=0)
    for k in ("status", "iters", "n_inliers", "E", "R", "t"):
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k
    assert np.array_equal(got["inliers"][:m].cpu().numpy(), want["inliers"])


def _edge_batch():
    """Ragged and degenerate match lists: tiny pairs, identical points, collinear points, pure noise, a pair at the maximum
    list length of the path (4 800), plus unused capacity behind the last pair."""
    from oracle.gen_golden_pose import scene
    rng = np.random.default_rng(17)
    parts = []
    for m, outl in ((5, 0.0), (6, 0.0), (9, 0.3), (31, 0.2), (33, 0.5), (4800, 0.3)):
        s = scene(rng, m, outl, 0.1, 600.0, 640.0)
        parts.append((s[0], s[1], s[2], s[3]))
    K = parts[0][2]
    same = np.full((40, 2), 123.5, dtype=np.float32)
    parts.append((same, same.copy(), K, K))                                         # all matches identical
    line = np.stack([np.linspace(10, 500, 50), np.linspace(20, 400, 50)], 1).astype(np.float32)
    parts.append((line, line[::-1].copy(), K, K))                                   # collinear in both images
    parts.append((rng.uniform(0, 600, (200, 2)).astype(np.float32), rng.uniform(0, 600, (200, 2)).astype(np.float32), K, K))
    parts.append((np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32), K, K))
    mk0 = np.concatenate([p[0] for p in parts] + [np.full((100, 2), 7.0, np.float32)])   # 100 rows of unused capacity
    mk1 = np.concatenate([p[1] for p in parts] + [np.full((100, 2), 9.0, np.float32)])
    counts = np.array([len(p[0]) for p in parts], dtype=np.int32)
    return mk0, mk1, counts, np.stack([p[2] for p in parts]), np.stack([p[3] for p in parts])


@pytest.mark.gpu
def test_gpu_pose_edge_cases_against_oracle():
    from pope_b200 import pose
    mk0, mk1, counts, K0, K1 = _edge_batch()
    dev = torch.device("cuda:0")
    for conf, max_iters in ((0.99999, 200), (0.99, 1)):
        got = pose.estimate_pose_batch(torch.from_numpy(mk0).to(dev), torch.from_numpy(mk1).to(dev),
                                       torch.from_numpy(counts).to(dev), torch.from_numpy(K0), torch.from_numpy(K1), 0.5, conf,
                                       max_iters, seed=5)
        m = int(counts.sum())
        want = O.estimate_pose_batch(mk0[:m], mk1[:m], counts, K0, K1, 0.5, conf, max_iters, seed=5)
        for k in ("status", "iters", "n_inliers", "E", "R", "t"):
            assert np.array_equal(got[k].cpu().numpy(), want[k], equal_nan=True), (k, conf)
        assert np.array_equal(got["inliers"][:m].cpu().numpy(), want["inliers"])
        assert not got["inliers"][m:].any()
        assert got["status"][-1].item() == 0 and got["status"][5].item() == 1


@pytest.mark.gpu
def test_gpu_pose_many_small_pairs_and_nonfinite_input():
    from oracle.gen_golden_pose import scene
    from pope_b200 import pose
    rng = np.random.default_rng(23)
    sc = [scene(rng, int(rng.integers(0, 40)), 0.2, 0.1, 600.0, 640.0) for _ in range(700)]
    mk0, mk1 = np.concatenate([s[0] for s in sc]), np.concatenate([s[1] for s in sc])
    counts = np.array([len(s[0]) for s in sc], dtype=np.int32)
    K0, K1 = np.stack([s[2] for s in sc]), np.stack([s[3] for s in sc])
    dev = torch.device("cuda:0")
    got = pose.estimate_pose_batch(torch.from_numpy(mk0).to(dev), torch.from_numpy(mk1).to(dev), torch.from_numpy(counts).to(dev),
                                   torch.from_numpy(K0), torch.from_numpy(K1), 0.5, 0.99, 64, seed=1)
    sel = list(range(0, 700, 23))                      # the oracle is slow: check a spread of pairs exactly
    off = np.concatenate([[0], np.cumsum(counts)])
    for p in sel:
        w = O.estimate_pose(mk0[off[p]:off[p + 1]], mk1[off[p]:off[p + 1]], K0[p], K1[p], 0.5, 0.99, 64, 1, p)
        assert got["status"][p].item() == w["status"] and got["iters"][p].item() == w["iters"]
        assert np.array_equal(got["R"][p].cpu().numpy(), w["R"]) and np.array_equal(got["t"][p].cpu().numpy(), w["t"])
        assert np.array_equal(got["inliers"][off[p]:off[p + 1]].cpu().numpy(), w["inliers"])
    assert bool(((got["status"] == 1) == (got["n_inliers"] > 0)).all())
    # counts that promise more matches than the arrays hold are clipped to the capacity, not followed out of bounds
    lying = counts.copy()
    lying[-1] += 10_000
    got2 = pose.estimate_pose_batch(torch.from_numpy(mk0).to(dev), torch.from_numpy(mk1).to(dev), torch.from_numpy(lying).to(dev),
                                    torch.from_numpy(K0), torch.from_numpy(K1), 0.5, 0.99, 64, seed=1)
    assert torch.equal(got2["R"][:-1], got["R"][:-1]) and torch.equal(got2["inliers"], got["inliers"])
    # a batch without a single match is valid input (every pair `None`)
    none = pose.estimate_pose_batch(torch.zeros(0, 2, device=dev), torch.zeros(0, 2, device=dev),
                                    torch.zeros(3, dtype=torch.int32, device=dev), torch.from_numpy(K0[:3]), torch.from_numpy(K1[:3]),
                                    0.5, 0.99)
    assert none["status"].tolist() == [0, 0, 0] and none["inliers"].numel() == 0
    # non-finite coordinates must neither hang nor produce a pose for the poisoned pair
    bad0 = mk0.copy()
    p = int(np.argmax(counts))
    bad0[off[p]:off[p + 1]] = np.nan
    got = pose.estimate_pose_batch(torch.from_numpy(bad0).to(dev), torch.from_numpy(mk1).to(dev), torch.from_numpy(counts).to(dev),
                                   torch.from_numpy(K0), torch.from_numpy(K1), 0.5, 0.99, 64, seed=1)
    torch.cuda.synchronize()
    assert got["status"][p].item() == 0 and got["n_inliers"][p].item() == 0


@pytest.mark.gpu
def test_compute_pose_errors_drop_in(gold):
    """metrics.py:97-133 through the batched device path, against the reference formulas applied to the oracle's poses;
    the match order is shuffled to cover callers whose m_bids are not sorted."""
    from types import SimpleNamespace
    from pope_b200 import pose
    dev = torch.device("cuda:0")
    n = len(gold["counts"])
    m_bids = np.repeat(np.arange(n), gold["counts"])
    perm = np.random.default_rng(0).permutation(len(m_bids))
    T = np.tile(np.eye(4), (n, 1, 1))
    T[:, :3, :3], T[:, :3, 3] = gold["R_gt"], gold["t_gt"]
    data = dict(m_bids=torch.from_numpy(m_bids[perm]).to(dev), mkpts0_f=torch.from_numpy(gold["mkpts0"][perm]).to(dev),
                mkpts1_f=torch.from_numpy(gold["mkpts1"][perm]).to(dev), K0=torch.from_numpy(gold["K0"]).to(dev),
                K1=torch.from_numpy(gold["K1"]).to(dev), T_0to1=torch.from_numpy(T).to(dev))
    cfg = SimpleNamespace(TRAINER=SimpleNamespace(RANSAC_PIXEL_THR=0.5, RANSAC_CONF=0.99999))
    pose.compute_pose_errors(data, cfg)
    # the stable sort restores the fixture's own order within each pair only up to the shuffle, so rebuild the oracle input
    order = np.argsort(m_bids[perm], kind="stable")
    want = O.estimate_pose_batch(gold["mkpts0"][perm][order], gold["mkpts1"][perm][order], gold["counts"], gold["K0"],
                                 gold["K1"], 0.5, 0.99999, 1000, seed=0)
    off = np.concatenate([[0], np.cumsum(gold["counts"])])
    assert len(data["R_errs"]) == len(data["t_errs"]) == len(data["inliers"]) == n
    for b in range(n):
        if not want["status"][b]:
            assert data["R_errs"][b] == np.inf and data["t_errs"][b] == np.inf and data["inliers"][b].size == 0
            continue
        R, t, t_gt = want["R"][b], want["t"][b], gold["t_gt"][b]
        t_err = np.rad2deg(np.arccos(np.clip(np.dot(t, t_gt) / (np.linalg.norm(t) * np.linalg.norm(t_gt)), -1.0, 1.0)))
        t_err = np.minimum(t_err, 180 - t_err)
        R_err = np.rad2deg(np.abs(np.arccos(np.clip((np.trace(np.dot(R.T, gold["R_gt"][b])) - 1) / 2, -1.0, 1.0))))
        assert abs(data["R_errs"][b] - R_err) < 1e-6 and abs(data["t_errs"][b] - t_err) < 1e-6
        inl_sorted = want["inliers"][off[b]:off[b + 1]]
        back = np.empty(len(m_bids), dtype=bool)
        back[order] = want["inliers"]
        assert np.array_equal(data["inliers"][b], back[m_bids[perm] == b]) and data["inliers"][b].sum() == inl_sorted.sum()


@pytest.mark.gpu
@pytest.mark.parametrize("hw_c,n,dtype", [((30, 40), 3, torch.bfloat16), ((60, 80), 4, torch.bfloat16), ((30, 40), 2, torch.float32)])
def test_full_chain_recovers_planted_pose(hw_c, n, dtype):
    """Round trip at the path's real sizes: plant a rigid motion in the synthetic features (synth.posed_pair_features),
    run match_pairs_device -> estimate_pose_batch on the device, get the motion back."""
    from pope_b200 import ops, pose, synth
    dev = torch.device("cuda:0")
    hc, wc = hw_c
    d = synth.posed_pair_features(41, n, hw_c=hw_c, dtype=dtype)
    res = ops.match_pairs_device(d["feat_c0"].to(dev), d["feat_c1"].to(dev), d["feat_f0"].to(dev), d["feat_f1"].to(dev),
                                 (hc * 8, wc * 8), hw_c, hw_c)
    K = d["K"].expand(n, 3, 3)
    got = pose.estimate_pose_batch(res["mkpts0_f"], res["mkpts1_f"], res["counts"], K, K, 1.0, 0.99999)
    m = res.total()
    planted = int((~torch.isnan(d["proj"][..., 0])).sum())
    assert m > 0.6 * planted
    proj = d["proj"][res["b_ids"][:m].cpu(), res["i_ids"][:m].cpu()]
    resid = (res["mkpts1_f"][:m].cpu().double() - proj).norm(dim=1)
    assert int(resid.isnan().sum()) < 0.01 * m                   # (nearly) every match is a planted one ...
    assert float(resid[~resid.isnan()].median()) < 1.0           # ... and FineMatching lands within the 2 px fine grid
    assert bool((got["status"][:n] == 1).all())
    # the winning minimal-sample model is not refined (as in OpenCV) and the matches carry about a pixel of quantisation
    # noise, so single pairs may be a few degrees off; the strict check is the oracle equality below
    r_errs = [rot_angle(got["R"][p].cpu().numpy(), d["R"][p].numpy()) for p in range(n)]
    r_max, r_med = (2.5, 1.0) if hw_c == (60, 80) else (6.0, 3.0)      # 320x240 at f = 500 is a narrow, ill-conditioned view
    assert max(r_errs) < r_max and np.median(r_errs) < r_med, r_errs
    for p in range(n):
        assert got["n_inliers"][p].item() > 0.7 * res["counts"][p].item()
    if hw_c == (30, 40):                                         # and the pose stage equals its oracle on these lists
        counts = res["counts"][:n].cpu().numpy()
        want = O.estimate_pose_batch(res["mkpts0_f"][:m].cpu().numpy(), res["mkpts1_f"][:m].cpu().numpy(), counts, K.numpy(),
                                     K.numpy(), 1.0, 0.99999, 1000, seed=0)
        for k in ("status", "iters", "n_inliers", "R", "t"):
            assert np.array_equal(got[k].cpu().numpy(), want[k]), k


@pytest.mark.gpu
def test_pose_call_is_graph_capturable(gold):
    """No host synchronisation and no allocation outside torch's pool inside the call: it can be captured once and replayed."""
    from pope_b200 import pose
    dev = torch.device("cuda:0")
    args = [torch.from_numpy(gold[k]).to(dev) for k in ("mkpts0", "mkpts1", "counts", "K0", "K1")]
    eager = pose.estimate_pose_batch(*args, 0.5, 0.99999)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = pose.estimate_pose_batch(*args, 0.5, 0.99999, workspace=eager["workspace"])
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    for k in ("status", "iters", "n_inliers", "inliers", "R", "t", "E"):
        assert torch.equal(out[k], eager[k]), k
