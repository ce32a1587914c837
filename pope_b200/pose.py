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
