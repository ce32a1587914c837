"""CoarseMatching -- drop-in for src/matcher/utils/coarse_matching.py:59-261, computed by libpope_b200.so.

Same constructor config keys, same call signature `forward(feat_c0, feat_c1, data, mask_c0=None, mask_c1=None)`,
same keys written into `data` (b_ids, i_ids, j_ids, gt_mask, m_bids, mkpts0_c, mkpts1_c, mconf) with the same
dtypes and (b, i) ordering.  Deviations, all documented in DESIGN.md:
  * `data['conf_matrix']` is not produced (the L x S matrix never exists; only the training loss reads it) unless the
    debug switch `materialize_conf_matrix` is set: then it is written with the reference's own torch op sequence on the
    device (an extra output for inspection; the match list still comes from the CUDA path and never reads it).
  * padding masks (`mask_c0/mask_c1`, `data['mask0']`; :115-118, :28-43, :180-182) are handled by running the CUDA path on
    the valid cells of every pair (an invalid cell has similarity -1e9 in the reference, i.e. weight exactly 0 in both
    softmaxes and no match) and clearing the padded border afterwards -- see `_forward_masked`.
  * the training-time sampling branch (:200-236) is not implemented -> NotImplementedError.
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
        self.materialize_conf_matrix = False     # debug: also write data['conf_matrix'] (92 MB per 480x640 pair)
        self._workspace = None

    @torch.no_grad()
    def forward(self, feat_c0, feat_c1, data, mask_c0=None, mask_c1=None):
        if self.training:
            raise NotImplementedError("the CUDA coarse matcher is inference-only (call .eval())")
        if self.materialize_conf_matrix:          # coarse_matching.py:106-119, verbatim op sequence, inspection only
            c = feat_c0.shape[-1]
            sim = torch.einsum("nlc,nsc->nls", feat_c0.float() / c ** 0.5, feat_c1.float() / c ** 0.5) / self.temperature
            if mask_c0 is not None:
                sim.masked_fill_(~(mask_c0[..., None] * mask_c1[:, None]).bool(), -1e9)
            data["conf_matrix"] = torch.softmax(sim, 1) * torch.softmax(sim, 2)
        if mask_c0 is not None or mask_c1 is not None or "mask0" in data:
            return self._forward_masked(feat_c0, feat_c1, data, mask_c0, mask_c1)
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

    # ---- padded batches (coarse_matching.py:115-118 masked_fill(-INF), :28-43 mask_border_with_padding) -----------------
    @staticmethod
    def _interior_padded(mask2d: torch.Tensor, bd: int) -> torch.Tensor:
        """Cells of one image that survive `mask_border_with_padding`: the first bd rows / columns and everything from
        (valid height - bd) / (valid width - bd) on are cleared; valid extents and slice semantics as in the reference."""
        keep = torch.ones_like(mask2d, dtype=torch.bool)
        if bd > 0:
            hv, wv = int(mask2d.sum(0).max()), int(mask2d.sum(1).max())
            keep[:bd] = False
            keep[:, :bd] = False
            keep[hv - bd:] = False
            keep[:, wv - bd:] = False
        return keep.reshape(-1)

    def _forward_masked(self, feat_c0, feat_c1, data, mask_c0, mask_c1):
        """An invalid cell's similarities are -1e9 in the reference: exp() of them is exactly 0 in both softmaxes, so the
        valid cells see the same sums as if the invalid ones did not exist, and an invalid cell never matches (its own
        confidences are 0 or 1/(L S)).  Each pair therefore runs through the CUDA path on its valid cells only (gathered
        into a dense list, no border inside the kernel); the ids are mapped back and the border of the padded grid is
        cleared afterwards -- equivalent, because border removal only clears threshold-mask cells and never changes the
        row / column maxima of the mutual test (:176-189)."""
        n, L, _ = feat_c0.shape
        S = feat_c1.shape[1]
        hw0_c, hw1_c = tuple(data["hw0_c"]), tuple(data["hw1_c"])
        dev = feat_c0.device
        # like the reference, the similarity matrix is masked only when mask_c0 is passed (:115-118); data['mask0'] alone
        # selects the padded border handling (:180-182) and nothing else
        fill = mask_c0 is not None
        if fill:
            m0 = mask_c0.reshape(n, L).bool()
            m1 = (mask_c1 if mask_c1 is not None else torch.ones(n, S, dtype=torch.bool, device=dev)).reshape(n, S).bool()
        all0, all1 = torch.arange(L, device=dev), torch.arange(S, device=dev)
        scale = data["hw0_i"][0] / data["hw0_c"][0]
        parts = {k: [] for k in ("b_ids", "i_ids", "j_ids", "mconf")}
        for b in range(n):
            v0 = torch.nonzero(m0[b]).reshape(-1) if fill else all0
            v1 = torch.nonzero(m1[b]).reshape(-1) if fill else all1
            if v0.numel() == 0 or v1.numel() == 0:
                continue
            res = ops.coarse_match(feat_c0[b, v0][None].contiguous(), feat_c1[b, v1][None].contiguous(), (v0.numel(), 1),
                                   (v1.numel(), 1), pixel_scale=scale, thr=self.thr, border_rm=0,
                                   temperature=self.temperature, impl=self.impl, workspace=self._workspace)
            self._workspace = res["workspace"]        # sized for the largest pair so far, reused for the others
            out = res.sliced()
            i_ids, j_ids = v0[out["i_ids"]], v1[out["j_ids"]]
            if "mask0" in data:
                keep0 = self._interior_padded(data["mask0"][b].bool(), self.border_rm)
                keep1 = self._interior_padded(data["mask1"][b].bool(), self.border_rm)
            else:                                   # masks given to forward() only: plain mask_border on the full grids
                keep0 = self._interior_padded(torch.ones(hw0_c, dtype=torch.bool, device=dev), self.border_rm)
                keep1 = self._interior_padded(torch.ones(hw1_c, dtype=torch.bool, device=dev), self.border_rm)
            keep = keep0.to(dev)[i_ids] & keep1.to(dev)[j_ids]
            parts["b_ids"].append(torch.full_like(i_ids[keep], b))
            parts["i_ids"].append(i_ids[keep])
            parts["j_ids"].append(j_ids[keep])
            parts["mconf"].append(out["mconf"][keep])
        cat = lambda k, dt: (torch.cat(parts[k]) if parts[k] else torch.empty(0, dtype=dt, device=dev))
        b_ids, i_ids, j_ids, mconf = cat("b_ids", torch.int64), cat("i_ids", torch.int64), cat("j_ids", torch.int64), \
            cat("mconf", torch.float32)
        mk0 = torch.stack([i_ids % hw0_c[1], i_ids // hw0_c[1]], dim=1) * scale
        mk1 = torch.stack([j_ids % hw1_c[1], j_ids // hw1_c[1]], dim=1) * scale
        if "scale0" in data:
            mk0 = mk0 * data["scale0"][b_ids]
            mk1 = mk1 * data["scale1"][b_ids]
        data.update(b_ids=b_ids, i_ids=i_ids, j_ids=j_ids, gt_mask=mconf == 0, m_bids=b_ids, mkpts0_c=mk0, mkpts1_c=mk1,
                    mconf=mconf)
