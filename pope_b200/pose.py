"""Relative pose from match lists on the device: the batched replacement of `estimate_pose` of the reference
(src/utils/metrics.py:69-94 -- cv2.findEssentialMat(..., method=cv2.RANSAC) + cv2.recoverPose, one CPU call per pair,
used at eval_onepose_json.py:164, acc1-30_onepose.py:152, visual_3dbbox.py:117 and metrics.py:119).

`estimate_pose_batch` takes the match lists where the Matcher hot path leaves them (device mkpts0_f / mkpts1_f, per-pair
counts) and solves all pairs in one set of launches (csrc/pose.cu); `estimate_pose` keeps the reference's per-pair signature
and return value for call sites that have numpy arrays.  There is no CPU path: both need the CUDA library and a device.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, require_cuda, stream_ptr

MAX_ITERS = 1024        # POPE_POSE_MAX_ITERS; OpenCV's own default bound is 1000


def estimate_pose_batch(mkpts0: torch.Tensor, mkpts1: torch.Tensor, counts: torch.Tensor, K0: torch.Tensor, K1: torch.Tensor,
                        thresh: float, conf: float = 0.99999, max_iters: int = 1000, seed: int = 0,
                        workspace: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """mkpts0 / mkpts1: float32 [cap, 2] pixel coordinates (device), the matches of pair p at rows
    [sum(counts[:p]), sum(counts[:p+1])); counts: int32 [n] (device; extra trailing entries are ignored when `K0` has
    fewer rows); K0 / K1: [n, 3, 3] intrinsics.  Returns device tensors: R float64 [n,3,3], t float64 [n,3], E float64
    [n,3,3], inliers bool [cap], n_inliers / status / iters int32 [n].  status 0 is the reference's `return None`."""
    dev = require_cuda(mkpts0, mkpts1, counts)
    n = int(K0.shape[0])
    if mkpts0.dtype != torch.float32 or mkpts1.dtype != torch.float32 or counts.dtype != torch.int32:
        raise _lib.PopeError("mkpts0 / mkpts1 must be float32 and counts int32")
    if mkpts0.shape != mkpts1.shape or mkpts0.dim() != 2 or mkpts0.shape[1] != 2 or counts.numel() < n or K1.shape[0] != n:
        raise _lib.PopeError("shape mismatch between mkpts0 / mkpts1 / counts / K0 / K1")
    mkpts0, mkpts1, counts = mkpts0.contiguous(), mkpts1.contiguous(), counts.contiguous()
    K0 = K0.to(device=dev, dtype=torch.float64).reshape(n, 9).contiguous()
    K1 = K1.to(device=dev, dtype=torch.float64).reshape(n, 9).contiguous()
    cap = int(mkpts0.shape[0])
    out = dict(R=torch.empty(n, 3, 3, dtype=torch.float64, device=dev), t=torch.empty(n, 3, dtype=torch.float64, device=dev),
               E=torch.empty(n, 3, 3, dtype=torch.float64, device=dev),
               inliers=torch.zeros(cap, dtype=torch.uint8, device=dev),
               n_inliers=torch.empty(n, dtype=torch.int32, device=dev), status=torch.empty(n, dtype=torch.int32, device=dev),
               iters=torch.empty(n, dtype=torch.int32, device=dev))
    if n == 0:
        out["inliers"] = out["inliers"].bool()
        return out
    h = lib()
    need = h.pope_pose_workspace_bytes(n, cap)
    if workspace is None or workspace.numel() < need or workspace.device != dev:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        st = h.pope_estimate_pose_batch(ptr(mkpts0), ptr(mkpts1), ptr(counts), n, cap, ptr(K0), ptr(K1), float(thresh),
                                        float(conf), int(max_iters), int(seed), ptr(out["R"]), ptr(out["t"]), ptr(out["E"]),
                                        ptr(out["inliers"]), ptr(out["n_inliers"]), ptr(out["status"]), ptr(out["iters"]),
                                        ptr(workspace), workspace.numel(), stream_ptr(dev))
    check(st, "pope_estimate_pose_batch")
    out["inliers"] = out["inliers"].bool()
    out["workspace"] = workspace
    return out


def estimate_pose(kpts0: np.ndarray, kpts1: np.ndarray, K0: np.ndarray, K1: np.ndarray, thresh: float, conf: float = 0.99999,
                  device: int = 0, seed: int = 0) -> Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]]:
    """Drop-in for src/utils/metrics.py:69 -- numpy in, `(R [3,3], t [3], inlier mask [M] bool)` or None out."""
    if len(kpts0) < 5:
        return None
    dev = torch.device("cuda", device)
    a = torch.as_tensor(np.ascontiguousarray(kpts0, dtype=np.float32)).to(dev)
    b = torch.as_tensor(np.ascontiguousarray(kpts1, dtype=np.float32)).to(dev)
    counts = torch.tensor([a.shape[0]], dtype=torch.int32, device=dev)
    out = estimate_pose_batch(a, b, counts, torch.as_tensor(np.asarray(K0, dtype=np.float64))[None],
                              torch.as_tensor(np.asarray(K1, dtype=np.float64))[None], thresh, conf, seed=seed)
    if int(out["status"][0].item()) == 0:
        return None
    return out["R"][0].cpu().numpy(), out["t"][0].cpu().numpy(), out["inliers"].cpu().numpy()


def relative_pose_error_batch(T_0to1: torch.Tensor, R: torch.Tensor, t: torch.Tensor, ignore_gt_t_thr: float = 0.0
                              ) -> Tuple[torch.Tensor, torch.Tensor]:
    """`relative_pose_error` of src/utils/metrics.py:10-24 for a batch, on the device of its inputs: T_0to1 [n,4,4],
    R [n,3,3], t [n,3] -> (t_err [n], R_err [n]) in degrees (float64)."""
    T = T_0to1.to(dtype=torch.float64, device=R.device)
    t_gt, R_gt = T[:, :3, 3], T[:, :3, :3]
    n = t.norm(dim=1) * t_gt.norm(dim=1)
    t_err = torch.rad2deg(torch.acos(torch.clamp((t * t_gt).sum(dim=1) / n, -1.0, 1.0)))
    t_err = torch.minimum(t_err, 180.0 - t_err)                   # handle E ambiguity
    t_err = torch.where(t_gt.norm(dim=1) < ignore_gt_t_thr, torch.zeros_like(t_err), t_err)
    cos = ((R.transpose(1, 2) @ R_gt).diagonal(dim1=1, dim2=2).sum(dim=1) - 1.0) / 2.0
    R_err = torch.rad2deg(torch.abs(torch.acos(torch.clamp(cos, -1.0, 1.0))))
    return t_err, R_err


def compute_pose_errors(data: dict, config, seed: int = 0) -> None:
    """Drop-in for `compute_pose_errors(data, config)` of src/utils/metrics.py:97-133 (the validation step's pose metrics):
    fills data['R_errs'], data['t_errs'] (lists of floats, inf where the reference's estimate_pose returns None) and
    data['inliers'] (list of bool arrays).  All pairs are solved in one batched call on the device that holds the matches;
    the only transfer is the result."""
    pixel_thr = config.TRAINER.RANSAC_PIXEL_THR
    conf = config.TRAINER.RANSAC_CONF
    m_bids, pts0, pts1 = data["m_bids"], data["mkpts0_f"], data["mkpts1_f"]
    n = int(data["K0"].shape[0])
    order = torch.sort(m_bids, stable=True).indices           # the Matcher emits sorted ids; other callers may not
    counts = torch.bincount(m_bids, minlength=n).to(torch.int32)
    out = estimate_pose_batch(pts0[order].float(), pts1[order].float(), counts, data["K0"], data["K1"], pixel_thr, conf,
                              seed=seed)
    t_err, R_err = relative_pose_error_batch(data["T_0to1"], out["R"], out["t"], ignore_gt_t_thr=0.0)
    ok = out["status"] == 1
    inf = torch.full_like(t_err, float("inf"))
    R_errs, t_errs = torch.where(ok, R_err, inf).cpu().tolist(), torch.where(ok, t_err, inf).cpu().tolist()
    inl = torch.empty_like(out["inliers"])
    inl[order] = out["inliers"]                                # back to the caller's match order
    inl, m_host, ok_host = inl.cpu().numpy(), m_bids.cpu().numpy(), ok.cpu().tolist()
    data.update({"R_errs": R_errs, "t_errs": t_errs,
                 "inliers": [inl[m_host == b] if ok_host[b] else np.array([]).astype(bool) for b in range(n)]})
