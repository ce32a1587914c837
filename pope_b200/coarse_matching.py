"""CoarseMatching -- drop-in for src/matcher/utils/coarse_matching.py:59-261, computed by libpope_b200.so.

Same constructor config keys, same call signature `forward(feat_c0, feat_c1, data, mask_c0=None, mask_c1=None)`,
same keys written into `data` (b_ids, i_ids, j_ids, gt_mask, m_bids, mkpts0_c, mkpts1_c, mconf) with the same
dtypes and (b, i) ordering.  Deviations, all documented in DESIGN.md:
  * `data['conf_matrix']` is not produced (the L x S matrix never exists; only the training loss reads it).
  * padding masks (`mask_c0/mask_c1`, `data['mask0']`) and the training-time sampling branch (:200-236) are not
    implemented on the CUDA path -> NotImplementedError (no inference caller of POPE passes them).
  * match_type 'sinkhorn' needs an unshipped superglue.py in the reference itself (:75-78) -> NotImplementedError.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops


class CoarseMatching(nn.Module):
    def __init__(self, config: dict):
        super().__init__()
        self.config = config
        self.thr = config["thr"]
        self.border_rm = config["border_rm"]
        self.train_coarse_percent = config["train_coarse_percent"]
        self.train_pad_num_gt_min = config["train_pad_num_gt_min"]
        self.match_type = config["match_type"]
        if self.match_type != "dual_softmax":
            raise NotImplementedError("only match_type='dual_softmax' is implemented (the reference's sinkhorn branch "
                                      "imports a superglue.py it does not ship)")
        self.temperature = config["dsmax_temperature"]
        self.impl = _lib.COARSE_AUTO      # POPE_COARSE_AUTO | _SIMT | _TCGEN05
        self._workspace = None

    @torch.no_grad()
    def forward(self, feat_c0, feat_c1, data, mask_c0=None, mask_c1=None):
        if mask_c0 is not None or mask_c1 is not None or "mask0" in data:
            raise NotImplementedError("padding masks are a training/MegaDepth feature not supported by the CUDA path")
        if self.training:
            raise NotImplementedError("the CUDA coarse matcher is inference-only (call .eval())")
        res = ops.coarse_match(feat_c0, feat_c1, data["hw0_c"], data["hw1_c"],
                               pixel_scale=data["hw0_i"][0] / data["hw0_c"][0], thr=self.thr,
                               border_rm=self.border_rm, temperature=self.temperature, impl=self.impl,
                               workspace=self._workspace)
        self._workspace = res["workspace"]
        out = res.sliced()                      # the single host sync of the coarse stage (reads M)
        if "scale0" in data:                    # per-image rescale of coarse_matching.py:243-250
            s = data["hw0_i"][0] / data["hw0_c"][0]
            out["mkpts0_c"] = out["mkpts0_c"] / s * (s * data["scale0"][out["b_ids"]])
            out["mkpts1_c"] = out["mkpts1_c"] / s * (s * data["scale1"][out["b_ids"]])
        data.update(**out)
