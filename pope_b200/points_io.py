"""On-disk match format of the reference, written by libpope_b200.so.

The reference stores every pair's matches with `np.savetxt` (linemod.py:147-171: `data/<set>-points/<object>/{pre_bbox,
mkpts0,mkpts1,pre_K}/<pair>.txt`) and `pose/dataset.py` reads them back with `np.loadtxt`.  `savetxt` here produces
byte-identical files; `write_match_files` writes the mkpts0 / mkpts1 files of a whole batch from the host pipeline's
output slots on a pool of native threads (no per-pair Python loop, no per-value Python string formatting)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from ._lib import PopeError, check, lib


def savetxt(path: str, a) -> None:
    """np.savetxt(path, a) for a float32 / float64 array of 1 or 2 dimensions (default format, delimiter, newline)."""
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    if a.ndim == 0 or a.ndim > 2:
        raise PopeError("savetxt takes 1-D or 2-D arrays (like numpy.savetxt)")
    rows, cols = (a.shape[0], 1) if a.ndim == 1 else a.shape
    if cols == 0:
        raise PopeError("savetxt: arrays without columns are not supported")
    if a.dtype == np.float32:
        a = np.ascontiguousarray(a)
        st = lib().pope_savetxt_f32(os.fsencode(path), a.ctypes.data, rows, cols)
    else:
        a = np.ascontiguousarray(a, dtype=np.float64)
        st = lib().pope_savetxt_f64(os.fsencode(path), a.ctypes.data, rows, cols)
    check(st, "pope_savetxt")


def write_match_files(out_dir: str, names: Sequence[str], out: Dict[str, torch.Tensor], min_matches: int = 5,
                      threads: int = 0) -> int:
    """Writes <out_dir>/mkpts0/<name>.txt and <out_dir>/mkpts1/<name>.txt for every pair of a host-pipeline result
    (`Pipeline.run` / `match_pairs_host`: mkpts0_f, mkpts1_f [n, cap, 2] float32, counts [n] int32) that has at least
    `min_matches` matches (linemod.py:143-146 skips the others).  Returns the number of pairs written."""
    k0, k1, cnt = out["mkpts0_f"], out["mkpts1_f"], out["counts"]
    n, cap = k0.shape[0], k0.shape[1]
    if len(names) != n:
        raise PopeError("one file name per pair is required")
    for t in (k0, k1, cnt):
        if t.device.type != "cpu" or not t.is_contiguous():
            raise PopeError("write_match_files takes the contiguous host tensors of the pipeline")
    if k0.dtype != torch.float32 or k1.dtype != torch.float32 or cnt.dtype != torch.int32:
        raise PopeError("mkpts*_f must be float32 and counts int32")
    arr = (C.c_char_p * max(n, 1))(*[os.fsencode(s) for s in names])
    written = C.c_int32(0)
    st = lib().pope_write_match_files(os.fsencode(out_dir), arr, n, k0.data_ptr(), k1.data_ptr(), cnt.data_ptr(), cap,
                                      int(min_matches), int(threads), C.byref(written))
    check(st, "pope_write_match_files")
    return int(written.value)


def loadtxt(path: str) -> np.ndarray:
    """np.loadtxt(path, delimiter=' ') as pose/dataset.py:75-101 uses it: float64, a single row or a single column comes
    back 1-D, a single value 0-D, an empty file as an empty array."""
    size = os.path.getsize(path)
    buf = np.empty(size // 2 + 1, dtype=np.float64)              # a value takes at least two bytes of text
    rows, cols = C.c_int64(0), C.c_int(0)
    st = lib().pope_loadtxt_f64(os.fsencode(path), buf.ctypes.data, buf.size, C.byref(rows), C.byref(cols))
    check(st, "pope_loadtxt_f64")
    a = buf[: rows.value * cols.value].reshape(rows.value, cols.value).copy()
    return np.squeeze(a) if a.size else np.empty(0, dtype=np.float64)


def read_match_files(in_dir: str, names: Sequence[str], capacity: int, threads: int = 0) -> Dict[str, torch.Tensor]:
    """The inverse of `write_match_files`: per-pair slots mkpts0_f / mkpts1_f [n, capacity, 2] float32 and counts [n] int32
    (-1 where the pair has no files), pinned when CUDA is available so that they can go straight to the device."""
    n = len(names)
    pin = torch.cuda.is_available()
    k0 = torch.zeros(n, capacity, 2, dtype=torch.float32, pin_memory=pin)
    k1 = torch.zeros(n, capacity, 2, dtype=torch.float32, pin_memory=pin)
    cnt = torch.zeros(n, dtype=torch.int32, pin_memory=pin)
    arr = (C.c_char_p * max(n, 1))(*[os.fsencode(s) for s in names])
    st = lib().pope_read_match_files(os.fsencode(in_dir), arr, n, k0.data_ptr(), k1.data_ptr(), cnt.data_ptr(), int(capacity),
                                     int(threads))
    check(st, "pope_read_match_files")
    return {"mkpts0_f": k0, "mkpts1_f": k1, "counts": cnt}


def imwrite_png(path: str, img, level: int = 1) -> None:
    """cv2.imwrite(path, img) for an 8-bit [h, w], [h, w, 1], [h, w, 3] (BGR) or [h, w, 4] (BGRA) array: the PNG decodes to
    exactly these pixels (linemod.py:172-173 / pose/dataset.py:102-103)."""
    a = img.detach().cpu().numpy() if torch.is_tensor(img) else np.asarray(img)
    if a.dtype != np.uint8 or a.ndim not in (2, 3):
        raise PopeError("imwrite_png takes uint8 arrays of shape [h, w] or [h, w, c]")
    a = np.ascontiguousarray(a)
    h, w = a.shape[:2]
    c = 1 if a.ndim == 2 else a.shape[2]
    check(lib().pope_write_png(os.fsencode(path), a.ctypes.data, h, w, c, w * c, int(level)), "pope_write_png")


def imwrite_png_batch(paths: Sequence[str], imgs: Sequence[np.ndarray], level: int = 1, threads: int = 0) -> None:
    """All crops of a batch on native threads; imgs: uint8 arrays with the same channel count, any sizes."""
    n = len(paths)
    if n != len(imgs):
        raise PopeError("one path per image is required")
    if n == 0:
        return
    arrs = [np.ascontiguousarray(i.detach().cpu().numpy() if torch.is_tensor(i) else i) for i in imgs]
    cs = {1 if a.ndim == 2 else a.shape[2] for a in arrs}
    if len(cs) != 1 or any(a.dtype != np.uint8 or a.ndim not in (2, 3) for a in arrs):
        raise PopeError("imwrite_png_batch takes uint8 images with one common channel count")
    p = (C.c_char_p * n)(*[os.fsencode(s) for s in paths])
    d = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    hs = np.array([a.shape[0] for a in arrs], dtype=np.int32)
    ws = np.array([a.shape[1] for a in arrs], dtype=np.int32)
    check(lib().pope_write_png_batch(p, d, hs.ctypes.data, ws.ctypes.data, cs.pop(), n, int(level), int(threads)),
          "pope_write_png_batch")


def write_pair_records(out_dir: str, names: Sequence[str], out: Dict[str, torch.Tensor], pre_bbox, pre_K, crops0=None,
                       crops1=None, min_matches: int = 5, threads: int = 0) -> int:
    """Everything linemod.py:147-173 stores for the pairs of one object directory, for a whole batch:
    <out_dir>/{pre_bbox,mkpts0,mkpts1,pre_K}/<name>.txt and, when crops are given, <out_dir>/{img0,img1}/<name>.png.
    `out`: the host pipeline's per-pair slots (see write_match_files); pre_bbox [n, 4], pre_K [n, 3, 3]; crops0 / crops1:
    lists of uint8 BGR images.  Pairs with fewer than `min_matches` matches or a pre_K that is not 3 x 3 are skipped like the
    reference does (:142-146).  Returns the number of pairs written."""
    n = len(names)
    counts = out["counts"].tolist()
    pre_bbox, pre_K = np.asarray(pre_bbox, dtype=np.float64), np.asarray(pre_K, dtype=np.float64)
    if pre_bbox.shape[0] != n or pre_K.shape[0] != n or pre_K.shape[1:] != (3, 3):
        raise PopeError("pre_bbox [n, 4] and pre_K [n, 3, 3] are required, one per pair")
    for sub in ("pre_bbox", "pre_K") + (("img0", "img1") if crops0 is not None else ()):
        os.makedirs(os.path.join(out_dir, sub), exist_ok=True)
    written = write_match_files(out_dir, names, out, min_matches, threads)
    live = [p for p in range(n) if counts[p] >= min_matches]
    for p in live:                                               # two tiny files per pair
        savetxt(os.path.join(out_dir, "pre_bbox", names[p] + ".txt"), pre_bbox[p])
        savetxt(os.path.join(out_dir, "pre_K", names[p] + ".txt"), pre_K[p])
    if crops0 is not None:
        imwrite_png_batch([os.path.join(out_dir, "img0", names[p] + ".png") for p in live] +
                          [os.path.join(out_dir, "img1", names[p] + ".png") for p in live],
                          [crops0[p] for p in live] + [crops1[p] for p in live], threads=threads)
    return written
