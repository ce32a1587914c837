"""On-disk match format: the native writer must produce the bytes numpy.savetxt produces --
numpy.savetxt IS the reference here (linemod.py:168-171) -- and numpy.loadtxt (pose/dataset.py) must read them back.
Host code only: runs without a GPU (the library loads without one)."""
import os

import numpy as np
import pytest
import torch

from pope_b200 import points_io


def _same_bytes(tmp_path, a, name):
    ref, got = tmp_path / f"{name}_np.txt", tmp_path / f"{name}_b200.txt"
    np.savetxt(ref, a)
    points_io.savetxt(str(got), a)
    assert got.read_bytes() == ref.read_bytes(), name
    want = np.asarray(a)
    back = np.loadtxt(got).reshape(want.shape).astype(want.dtype)          # what pose/dataset.py reads
    assert np.array_equal(back, want, equal_nan=True), name


def test_savetxt_bytes_equal_numpy(tmp_path):
    rng = np.random.default_rng(0)
    mk = (rng.integers(0, 80, (257, 2)) * 8 + rng.uniform(-4, 4, (257, 2))).astype(np.float32)     # refined keypoints
    _same_bytes(tmp_path, mk, "mkpts")
    _same_bytes(tmp_path, torch.from_numpy(mk), "mkpts_torch")
    _same_bytes(tmp_path, np.array([12.0, 40.5, 300.25, 411.0]), "pre_bbox_1d")                     # 1-D: one value per line
    _same_bytes(tmp_path, np.array([[572.4114, 0.0, 325.2611], [0.0, 573.57043, 242.04899], [0.0, 0.0, 1.0]]), "pre_K")
    special = np.array([[0.0, -0.0], [1e-45, -3.4028235e38], [np.inf, -np.inf], [np.nan, 1.17549435e-38]], dtype=np.float32)
    _same_bytes(tmp_path, special, "special")
    _same_bytes(tmp_path, rng.standard_normal((1000, 2)), "f64")
    _same_bytes(tmp_path, rng.standard_normal((3, 7)).astype(np.float32), "wide")


def test_write_match_files_batch(tmp_path):
    """Batch writer from the pipeline's per-pair slots: one pair below the 5-match limit is skipped, the files of the
    others equal np.savetxt of the live rows."""
    rng = np.random.default_rng(1)
    n, cap = 6, 40
    counts = torch.tensor([12, 4, 40, 0, 5, 33], dtype=torch.int32)
    out = {"mkpts0_f": torch.from_numpy(rng.uniform(0, 640, (n, cap, 2)).astype(np.float32)),
           "mkpts1_f": torch.from_numpy(rng.uniform(0, 640, (n, cap, 2)).astype(np.float32)), "counts": counts}
    names = [f"{p:04d}.png-{p + 7:04d}.png" for p in range(n)]
    written = points_io.write_match_files(str(tmp_path / "pts"), names, out, min_matches=5, threads=3)
    assert written == 4
    for p in range(n):
        for key, sub in (("mkpts0_f", "mkpts0"), ("mkpts1_f", "mkpts1")):
            path = tmp_path / "pts" / sub / f"{names[p]}.txt"
            if counts[p] < 5:
                assert not path.exists()
                continue
            ref = tmp_path / "ref.txt"
            np.savetxt(ref, out[key][p, : counts[p]].numpy())
            assert path.read_bytes() == ref.read_bytes()
            assert np.loadtxt(path).shape == (int(counts[p]), 2)


def test_points_io_errors(tmp_path):
    from pope_b200._lib import PopeError
    with pytest.raises(PopeError):
        points_io.savetxt(str(tmp_path / "no_such_dir" / "x.txt"), np.zeros((2, 2), dtype=np.float32))
    with pytest.raises(PopeError):
        points_io.savetxt(str(tmp_path / "x.txt"), np.zeros((2, 2, 2)))


def _same_array(tmp_path, a, name, text=None):
    """loadtxt must return what numpy.loadtxt(path, delimiter=' ') returns (pose/dataset.py:75-101)."""
    path = tmp_path / f"{name}.txt"
    if text is None:
        np.savetxt(path, a)
    else:
        path.write_text(text)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = np.loadtxt(path, delimiter=" ") if text is None else np.loadtxt(path)
    got = points_io.loadtxt(str(path))
    assert got.dtype == np.float64 and got.shape == want.shape, (name, got.shape, want.shape)
    assert np.array_equal(got, want, equal_nan=True), name
    assert np.array_equal(np.signbit(got), np.signbit(want)), name


def test_loadtxt_equals_numpy(tmp_path):
    rng = np.random.default_rng(2)
    _same_array(tmp_path, (rng.uniform(0, 640, (300, 2))).astype(np.float32), "mkpts")
    _same_array(tmp_path, rng.standard_normal((500, 2)), "f64")
    _same_array(tmp_path, np.array([12.0, 40.5, 300.25, 411.0]), "bbox_1d")
    _same_array(tmp_path, np.array([[572.4114, 0.0, 325.2611], [0.0, 573.57043, 242.04899], [0.0, 0.0, 1.0]]), "K")
    _same_array(tmp_path, rng.standard_normal((1, 2)), "single_row")
    _same_array(tmp_path, np.array([3.25]), "single_value")
    _same_array(tmp_path, np.array([[0.0, -0.0], [1e-45, -3.4028235e38], [np.inf, -np.inf], [np.nan, 5e-324]]), "special")
    _same_array(tmp_path, None, "empty", text="")
    _same_array(tmp_path, None, "comments", text="# header\n1 2 3\n\n4 5 6  # tail\n")
    _same_array(tmp_path, None, "free_form", text="+1.5\t2e3   -7\r\n1e400 -1e400 1e-400\n")
    (tmp_path / "ragged.txt").write_text("1 2\n3\n")
    with pytest.raises(Exception):
        points_io.loadtxt(str(tmp_path / "ragged.txt"))
    (tmp_path / "word.txt").write_text("1 two\n")
    with pytest.raises(Exception):
        points_io.loadtxt(str(tmp_path / "word.txt"))
    with pytest.raises(Exception):
        points_io.loadtxt(str(tmp_path / "missing.txt"))


def test_read_match_files_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    n, cap = 7, 50
    out = {"mkpts0_f": torch.from_numpy(rng.uniform(0, 640, (n, cap, 2)).astype(np.float32)),
           "mkpts1_f": torch.from_numpy(rng.uniform(0, 480, (n, cap, 2)).astype(np.float32)),
           "counts": torch.tensor([50, 4, 17, 0, 5, 33, 9], dtype=torch.int32)}
    names = [f"{i:04d}-{i + 1:04d}" for i in range(n)]
    assert points_io.write_match_files(str(tmp_path / "obj"), names, out) == 5
    back = points_io.read_match_files(str(tmp_path / "obj"), names, cap, threads=3)
    assert back["counts"].tolist() == [50, -1, 17, -1, 5, 33, 9]
    for p, m in enumerate(back["counts"].tolist()):
        if m > 0:
            assert torch.equal(back["mkpts0_f"][p, :m], out["mkpts0_f"][p, :m])
            assert torch.equal(back["mkpts1_f"][p, :m], out["mkpts1_f"][p, :m])
    short = points_io.read_match_files(str(tmp_path / "obj"), names, 10)          # truncation at the slot capacity
    assert short["counts"].tolist() == [10, -1, 10, -1, 5, 10, 9]
    assert torch.equal(short["mkpts0_f"][0], out["mkpts0_f"][0, :10])


def test_png_crops_decode_to_the_input(tmp_path):
    """linemod.py:172-173 writes the two crops with cv2.imwrite, pose/dataset.py:102-103 reads them with cv2.imread: the
    native writer's files must decode to the same pixels as cv2's own files (and as the input)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(4)
    yy, xx = np.mgrid[0:97, 0:131]
    smooth = np.stack([128 + 100 * np.sin(xx / 17.0 + yy / 29.0), 128 + 90 * np.cos(xx / 23.0), yy * 2.0], 2).clip(0, 255).astype(np.uint8)
    cases = [rng.integers(0, 256, (37, 53, 3), dtype=np.uint8), smooth, rng.integers(0, 256, (64, 64), dtype=np.uint8),
             rng.integers(0, 256, (20, 31, 1), dtype=np.uint8), rng.integers(0, 256, (15, 17, 4), dtype=np.uint8),
             np.zeros((1, 1, 3), np.uint8), np.full((5, 300, 3), 255, np.uint8)]
    for k, a in enumerate(cases):
        ours, ref = str(tmp_path / f"a{k}.png"), str(tmp_path / f"b{k}.png")
        points_io.imwrite_png(ours, a)
        assert cv2.imwrite(ref, a)
        for flag in (cv2.IMREAD_UNCHANGED, cv2.IMREAD_COLOR):
            x, y = cv2.imread(ours, flag), cv2.imread(ref, flag)
            assert x is not None and x.shape == y.shape and np.array_equal(x, y), (k, flag)
        assert np.array_equal(cv2.imread(ours, cv2.IMREAD_UNCHANGED).reshape(a.shape), a)
    paths = [str(tmp_path / f"batch{k}.png") for k in range(3)]
    points_io.imwrite_png_batch(paths, [cases[0], cases[1], torch.from_numpy(cases[6])], threads=2)
    for p, a in zip(paths, (cases[0], cases[1], cases[6])):
        assert np.array_equal(cv2.imread(p), a)
    with pytest.raises(Exception):
        points_io.imwrite_png(str(tmp_path / "bad.png"), np.zeros((4, 4, 2), np.uint8))
    with pytest.raises(Exception):
        points_io.imwrite_png(str(tmp_path / "bad.png"), np.zeros((4, 4, 3), np.float32))
    with pytest.raises(Exception):
        points_io.imwrite_png(str(tmp_path / "no_such_dir" / "x.png"), cases[0])


def test_write_pair_records_equals_the_reference_loop(tmp_path):
    """The six per-pair artefacts of linemod.py:147-173 for a batch, against that loop restated with numpy / cv2."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    n, cap = 5, 30
    out = {"mkpts0_f": torch.from_numpy(rng.uniform(0, 640, (n, cap, 2)).astype(np.float32)),
           "mkpts1_f": torch.from_numpy(rng.uniform(0, 480, (n, cap, 2)).astype(np.float32)),
           "counts": torch.tensor([30, 3, 12, 5, 0], dtype=torch.int32)}
    names = [f"{i:04d}-{i + 7:04d}" for i in range(n)]
    bbox = rng.uniform(0, 400, (n, 4))
    Ks = np.tile(np.array([[572.4, 0, 325.3], [0, 573.6, 242.0], [0, 0, 1.0]]), (n, 1, 1)) + rng.normal(0, 1, (n, 3, 3))
    c0 = [rng.integers(0, 256, (int(rng.integers(8, 40)), int(rng.integers(8, 40)), 3), dtype=np.uint8) for _ in range(n)]
    c1 = [rng.integers(0, 256, (int(rng.integers(8, 40)), int(rng.integers(8, 40)), 3), dtype=np.uint8) for _ in range(n)]
    got, ref = tmp_path / "got", tmp_path / "ref"
    assert points_io.write_pair_records(str(got), names, out, bbox, Ks, c0, c1) == 3
    for p in range(n):                                           # the reference loop (linemod.py:142-173)
        m = int(out["counts"][p])
        if m < 5:
            continue
        for sub in ("pre_bbox", "mkpts0", "mkpts1", "pre_K", "img0", "img1"):
            (ref / sub).mkdir(parents=True, exist_ok=True)
        np.savetxt(ref / "pre_bbox" / f"{names[p]}.txt", bbox[p])
        np.savetxt(ref / "mkpts0" / f"{names[p]}.txt", out["mkpts0_f"][p, :m].numpy())
        np.savetxt(ref / "mkpts1" / f"{names[p]}.txt", out["mkpts1_f"][p, :m].numpy())
        np.savetxt(ref / "pre_K" / f"{names[p]}.txt", Ks[p])
        cv2.imwrite(str(ref / "img0" / f"{names[p]}.png"), c0[p])
        cv2.imwrite(str(ref / "img1" / f"{names[p]}.png"), c1[p])
    for sub in ("pre_bbox", "mkpts0", "mkpts1", "pre_K", "img0", "img1"):
        want = sorted(os.listdir(ref / sub))
        assert sorted(os.listdir(got / sub)) == want and len(want) == 3, sub
        for f in want:
            if f.endswith(".txt"):
                assert (got / sub / f).read_bytes() == (ref / sub / f).read_bytes(), (sub, f)
            else:
                assert np.array_equal(cv2.imread(str(got / sub / f)), cv2.imread(str(ref / sub / f))), (sub, f)
