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
