"""FineMatching -- drop-in for src/matcher/utils/fine_matching.py:9-74, computed by libpope_b200.so
(`pope_fine_match`: one warp per match, correlation + softmax + expectation + std fused).

Writes `expec_f [M,3]`, `mkpts0_f`, `mkpts1_f` into `data`; the M == 0 branch (:33-41) is reproduced.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops


class FineMatching(nn.Module):
    def __init__(self):
        super().__init__()

    @torch.no_grad()
    def forward(self, feat_f0, feat_f1, data):
        M, WW, C = feat_f0.shape
        W = int(math.sqrt(WW))
        scale = data["hw0_i"][0] / data["hw0_f"][0]
        self.M, self.W, self.WW, self.C, self.scale = M, W, WW, C, scale
        if M == 0:
            assert not self.training, "M is always >0, when training, see coarse_matching.py"
            data.update({"expec_f": torch.empty(0, 3, device=feat_f0.device),
                         "mkpts0_f": data["mkpts0_c"], "mkpts1_f": data["mkpts1_c"]})
            return
        if feat_f0.dtype not in (torch.float32, torch.bfloat16):
            feat_f0, feat_f1 = feat_f0.float(), feat_f1.float()
        n_kept = len(data["mconf"])
        expec, mk1f = ops.fine_match(feat_f0, feat_f1, _padded(data["mkpts1_c"], M), (W // 2) * scale)
        data.update({"expec_f": expec})
        mkpts1_f = mk1f[:n_kept]
        if "scale0" in data:      # per-image rescale of fine_matching.py:68-69 (never used by POPE's callers)
            scale1 = scale * data["scale1"][data["b_ids"]]
            mkpts1_f = data["mkpts1_c"] + (expec[:, :2] * (W // 2) * scale1)[:n_kept]
        data.update({"mkpts0_f": data["mkpts0_c"], "mkpts1_f": mkpts1_f})


def _padded(mkpts1_c: torch.Tensor, M: int) -> torch.Tensor:
    """fine_matching.py:69 slices the refinement to len(mconf) (training pads M' > M); give the kernel M rows."""
    if mkpts1_c.shape[0] == M:
        return mkpts1_c
    out = mkpts1_c.new_zeros(M, 2)
    out[: mkpts1_c.shape[0]] = mkpts1_c
    return out
