"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference Matcher hot path (torch, fp32, CPU).

This is the *checker* for the CUDA path, never the product: only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.  The product package
(`pope_b200/`) does not import anything from `oracle/` and fails loudly without its CUDA library.

Pinning: the reference ships no golden vectors or tests for this path, so the
restatement is pinned against *outputs of the reference itself*: `oracle/gen_golden.py` imports the
unmodified reference (`/root/reference/src/matcher`, via `oracle/ref_shim.py`) in the build container,
runs its CoarseMatching / FinePreprocess / FineMatching on seeded inputs and commits the results under
`tests/golden/`;  `tests/test_oracle_golden.py` checks every function below against those fixtures
(bit-exact indices, fp32 values to 1e-6), and `tests/test_oracle_vs_reference.py` re-checks live
whenever /root/reference is present.

Every function cites the reference lines it follows (paths relative to /root/reference).
The op sequence is deliberately the same library-call sequence the reference performs on the CPU
(einsum -> softmax x softmax -> compare -> max -> where; unfold -> index; einsum -> softmax -> expectation),
so timing it is timing the reference's CPU path (bench.py `cpu_baseline.kind == "port"`).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# constants of src/matcher/utils/cvpr_ds_config.py:12-14,32-35,44
THR = 0.2
BORDER_RM = 2
DSMAX_TEMPERATURE = 0.1
FINE_WINDOW = 5


def dual_softmax_conf(feat_c0: torch.Tensor, feat_c1: torch.Tensor,
                      temperature: float = DSMAX_TEMPERATURE) -> torch.Tensor:
    """conf[n,l,s] = softmax_l(S)[n,l,s] * softmax_s(S)[n,l,s],  S = (f0/sqrt(C)) . (f1/sqrt(C)) / T.

    Follows src/matcher/utils/coarse_matching.py:106-119 (the dual_softmax branch, no padding mask)."""
    c = feat_c0.shape[-1]
    a = feat_c0 / c ** 0.5
    b = feat_c1 / c ** 0.5
    sim = torch.einsum("nlc,nsc->nls", a, b) / temperature
    return F.softmax(sim, 1) * F.softmax(sim, 2)


def _interior(h: int, w: int, bd: int) -> torch.Tensor:
    """Bool [h*w]: True for cells that survive the border removal of coarse_matching.py:8-25."""
    keep = torch.zeros(h, w, dtype=torch.bool)
    if bd <= 0:
        keep[:] = True
    elif h > 2 * bd and w > 2 * bd:
        keep[bd:h - bd, bd:w - bd] = True
    return keep.reshape(-1)


def coarse_match_from_conf(conf: torch.Tensor, hw0_i: Sequence[int], hw0_c: Sequence[int],
                           hw1_c: Sequence[int], thr: float = THR, border_rm: int = BORDER_RM
                           ) -> Dict[str, torch.Tensor]:
    """Threshold, border removal, mutual nearest neighbour, extraction, coarse pixel coordinates.

    Follows coarse_matching.py:167-196 and :239-259 (inference branch: no 'mask0', not training,
    no 'scale0').  The mutual test uses row/column maxima of the *full* matrix, border cells
    included (:187-189): only the threshold mask is border-cleared (:176-184)."""
    n, l, s = conf.shape
    keep0 = _interior(hw0_c[0], hw0_c[1], border_rm)
    keep1 = _interior(hw1_c[0], hw1_c[1], border_rm)
    sel = (conf > thr) & keep0[None, :, None] & keep1[None, None, :]
    sel &= conf == conf.max(dim=2, keepdim=True)[0]
    sel &= conf == conf.max(dim=1, keepdim=True)[0]
    # ":192 this only works when at most one True in each row"
    row_has, row_arg = sel.max(dim=2)
    b_ids, i_ids = torch.where(row_has)            # sorted by (b, i)
    j_ids = row_arg[b_ids, i_ids]
    mconf = conf[b_ids, i_ids, j_ids]
    scale = hw0_i[0] / hw0_c[0]
    mk0 = torch.stack([i_ids % hw0_c[1], i_ids // hw0_c[1]], dim=1) * scale
    mk1 = torch.stack([j_ids % hw1_c[1], j_ids // hw1_c[1]], dim=1) * scale
    live = mconf != 0
    return {
        "b_ids": b_ids, "i_ids": i_ids, "j_ids": j_ids,
        "gt_mask": mconf == 0, "m_bids": b_ids[live],
        "mkpts0_c": mk0[live], "mkpts1_c": mk1[live], "mconf": mconf[live],
    }


def coarse_match(feat_c0: torch.Tensor, feat_c1: torch.Tensor, hw0_i, hw0_c, hw1_c,
                 thr: float = THR, border_rm: int = BORDER_RM, temperature: float = DSMAX_TEMPERATURE,
                 chunk: int = 4) -> Dict[str, torch.Tensor]:
    """CoarseMatching.forward for a batch (coarse_matching.py:87-148), evaluated `chunk` pairs at a time
    so that 4800x4800 fp32 matrices (92 MB each, several live copies) fit in host memory.  Pairs are
    independent, so chunking changes nothing but `b_ids` offsets."""
    outs = []
    for b0 in range(0, feat_c0.shape[0], chunk):
        conf = dual_softmax_conf(feat_c0[b0:b0 + chunk].float(), feat_c1[b0:b0 + chunk].float(), temperature)
        o = coarse_match_from_conf(conf, hw0_i, hw0_c, hw1_c, thr, border_rm)
        o["b_ids"] = o["b_ids"] + b0
        o["m_bids"] = o["m_bids"] + b0
        outs.append(o)
    return {k: torch.cat([o[k] for o in outs]) for k in outs[0]}


def _interior_padded(mask2d: torch.Tensor, bd: int) -> torch.Tensor:
    """Bool [h*w]: cells that survive `mask_border_with_padding` (coarse_matching.py:28-43) for one image whose valid
    (unpadded) area is given by mask2d [h, w]: the first bd rows / columns and everything from (valid height - bd) /
    (valid width - bd) on are cleared, with the valid extents taken as the reference takes them (:38-39: the largest
    column sum and the largest row sum of the mask) and with Python's slice semantics for the upper bounds."""
    h, w = mask2d.shape
    keep = torch.ones(h, w, dtype=torch.bool)
    if bd <= 0:
        return keep.reshape(-1)
    hv = int(mask2d.sum(0).max())
    wv = int(mask2d.sum(1).max())
    keep[:bd] = False
    keep[:, :bd] = False
    keep[hv - bd:] = False
    keep[:, wv - bd:] = False
    return keep.reshape(-1)


def coarse_match_masked(feat_c0: torch.Tensor, feat_c1: torch.Tensor, hw0_i, hw0_c, hw1_c, mask0: torch.Tensor,
                        mask1: torch.Tensor, thr: float = THR, border_rm: int = BORDER_RM,
                        temperature: float = DSMAX_TEMPERATURE) -> Dict[str, torch.Tensor]:
    """CoarseMatching.forward with padding masks (the MegaDepth-style batches of the reference): mask0 [N, h0c, w0c],
    mask1 [N, h1c, w1c] bool, True = valid cell.  Follows coarse_matching.py:115-118 (similarities of invalid cells filled
    with -1e9 before the two softmaxes), :176-184 with `mask_border_with_padding` (:28-43) instead of `mask_border`,
    :187-196 and :239-259 as in `coarse_match_from_conf`."""
    n, l, c = feat_c0.shape
    s = feat_c1.shape[1]
    m0, m1 = mask0.reshape(n, l).bool(), mask1.reshape(n, s).bool()
    sim = torch.einsum("nlc,nsc->nls", feat_c0.float() / c ** 0.5, feat_c1.float() / c ** 0.5) / temperature
    sim = sim.masked_fill(~(m0[:, :, None] & m1[:, None, :]), -1e9)
    conf = F.softmax(sim, 1) * F.softmax(sim, 2)
    keep0 = torch.stack([_interior_padded(mask0[b].bool(), border_rm) for b in range(n)])
    keep1 = torch.stack([_interior_padded(mask1[b].bool(), border_rm) for b in range(n)])
    sel = (conf > thr) & keep0[:, :, None] & keep1[:, None, :]
    sel &= conf == conf.max(dim=2, keepdim=True)[0]
    sel &= conf == conf.max(dim=1, keepdim=True)[0]
    row_has, row_arg = sel.max(dim=2)
    b_ids, i_ids = torch.where(row_has)
    j_ids = row_arg[b_ids, i_ids]
    mconf = conf[b_ids, i_ids, j_ids]
    scale = hw0_i[0] / hw0_c[0]
    mk0 = torch.stack([i_ids % hw0_c[1], i_ids // hw0_c[1]], dim=1) * scale
    mk1 = torch.stack([j_ids % hw1_c[1], j_ids // hw1_c[1]], dim=1) * scale
    live = mconf != 0
    return {"b_ids": b_ids, "i_ids": i_ids, "j_ids": j_ids, "gt_mask": mconf == 0, "m_bids": b_ids[live],
            "mkpts0_c": mk0[live], "mkpts1_c": mk1[live], "mconf": mconf[live], "conf_matrix": conf}


def fine_windows(feat_f: torch.Tensor, b_ids: torch.Tensor, cell_ids: torch.Tensor, W: int = FINE_WINDOW,
                 stride: int = 4) -> torch.Tensor:
    """The unfold + gather of src/matcher/loftr_module/fine_preprocess.py:40-47 for one image:
    returns [M, W*W, C] with window element ww = ky*W + kx of coarse cell (y, x) equal to
    feat_f[b, :, stride*y - W//2 + ky, stride*x - W//2 + kx] (zero outside the map).

    Uses the reference's own op sequence (F.unfold then advanced indexing) so timings are comparable."""
    n, c, hf, wf = feat_f.shape
    if b_ids.numel() == 0:
        return torch.empty(0, W * W, c)
    cols = F.unfold(feat_f.float(), kernel_size=(W, W), stride=stride, padding=W // 2)   # [N, C*WW, Lc]
    cols = cols.reshape(n, c, W * W, -1).permute(0, 3, 2, 1)                            # n l ww c
    return cols[b_ids, cell_ids]


def fine_windows_direct(feat_f: torch.Tensor, b_ids: torch.Tensor, cell_ids: torch.Tensor, wc: int,
                        W: int = FINE_WINDOW, stride: int = 4) -> torch.Tensor:
    """Same result as `fine_windows` without materialising the 25x unfold (used to cross-check the
    restatement of the window geometry itself)."""
    n, c, hf, wf = feat_f.shape
    pad = W // 2
    padded = F.pad(feat_f.float(), (pad, pad, pad, pad))
    y = (cell_ids // wc) * stride
    x = (cell_ids % wc) * stride
    ky, kx = torch.meshgrid(torch.arange(W), torch.arange(W), indexing="ij")
    yy = y[:, None] + ky.reshape(-1)[None]
    xx = x[:, None] + kx.reshape(-1)[None]
    return padded[b_ids[:, None], :, yy, xx]                                              # [M, WW, C]


def fine_match(win0: torch.Tensor, win1: torch.Tensor, mkpts0_c: torch.Tensor, mkpts1_c: torch.Tensor,
               scale: float) -> Dict[str, torch.Tensor]:
    """FineMatching.forward + get_fine_match (src/matcher/utils/fine_matching.py:27-73), inference branch
    (no 'scale0').  `scale` = hw0_i[0] / hw0_f[0]."""
    m, ww, c = win0.shape
    w = int(math.sqrt(ww))
    if m == 0:
        return {"expec_f": torch.empty(0, 3), "mkpts0_f": mkpts0_c, "mkpts1_f": mkpts1_c}
    centre = win0[:, ww // 2, :].float()
    sim = torch.einsum("mc,mrc->mr", centre, win1.float())
    heat = torch.softmax(sim / c ** 0.5, dim=1)
    lin = torch.linspace(-1.0, 1.0, w)
    gx = lin.repeat(w)                      # x varies fastest
    gy = lin.repeat_interleave(w)
    ex = (heat * gx).sum(1)
    ey = (heat * gy).sum(1)
    var_x = (heat * gx ** 2).sum(1) - ex ** 2
    var_y = (heat * gy ** 2).sum(1) - ey ** 2
    std = torch.sqrt(var_x.clamp(min=1e-10)) + torch.sqrt(var_y.clamp(min=1e-10))
    coords = torch.stack([ex, ey], 1)
    return {
        "expec_f": torch.cat([coords, std[:, None]], 1),
        "mkpts0_f": mkpts0_c,
        "mkpts1_f": mkpts1_c + (coords * (w // 2) * scale)[: mkpts1_c.shape[0]],
    }


def linear_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """LinearAttention.forward without masks (src/matcher/loftr_module/linear_attention.py:21-47):
    q [N,L,H,D], k/v [N,S,H,D] -> [N,L,H,D]."""
    Q = F.elu(q) + 1
    K = F.elu(k) + 1
    s = v.size(1)
    KV = torch.einsum("nshd,nshv->nhdv", K, v / s)
    Z = 1 / (torch.einsum("nlhd,nhd->nlh", Q, K.sum(dim=1)) + eps)
    return torch.einsum("nlhd,nhdv,nlh->nlhv", Q, KV, Z) * s


def encoder_layer(x: torch.Tensor, source: torch.Tensor, w: Dict[str, torch.Tensor], nhead: int = 8) -> torch.Tensor:
    """LoFTREncoderLayer.forward (src/matcher/loftr_module/transformer.py:34-58).  `w` holds the layer's state dict
    (q_proj.weight, k_proj.weight, v_proj.weight, merge.weight, mlp.0.weight, mlp.2.weight, norm1/2.weight/bias)."""
    n, _, c = x.shape
    d = c // nhead
    q = F.linear(x, w["q_proj.weight"]).view(n, -1, nhead, d)
    k = F.linear(source, w["k_proj.weight"]).view(n, -1, nhead, d)
    v = F.linear(source, w["v_proj.weight"]).view(n, -1, nhead, d)
    msg = linear_attention(q, k, v).reshape(n, -1, c)
    msg = F.layer_norm(F.linear(msg, w["merge.weight"]), (c,), w["norm1.weight"], w["norm1.bias"])
    msg = F.linear(torch.relu(F.linear(torch.cat([x, msg], dim=2), w["mlp.0.weight"])), w["mlp.2.weight"])
    msg = F.layer_norm(msg, (c,), w["norm2.weight"], w["norm2.bias"])
    return x + msg


def fine_transformer(feat0: torch.Tensor, feat1: torch.Tensor, layers: Sequence[Dict[str, torch.Tensor]],
                     layer_names: Sequence[str], nhead: int = 8) -> Tuple[torch.Tensor, torch.Tensor]:
    """LocalFeatureTransformer.forward without masks (src/matcher/loftr_module/transformer.py:95-106):
    'self' updates each side from itself, 'cross' updates feat0 from feat1 and then feat1 from the NEW feat0."""
    feat0, feat1 = feat0.float(), feat1.float()
    for w, name in zip(layers, layer_names):
        if name == "self":
            feat0 = encoder_layer(feat0, feat0, w, nhead)
            feat1 = encoder_layer(feat1, feat1, w, nhead)
        elif name == "cross":
            feat0 = encoder_layer(feat0, feat1, w, nhead)
            feat1 = encoder_layer(feat1, feat0, w, nhead)
        else:
            raise KeyError(name)
    return feat0, feat1


def fine_merge_coarse(win0: torch.Tensor, win1: torch.Tensor, feat_c0: torch.Tensor, feat_c1: torch.Tensor,
                      b_ids: torch.Tensor, i_ids: torch.Tensor, j_ids: torch.Tensor, w: Dict[str, torch.Tensor]
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
    """The coarse-context branch of FinePreprocess.forward (src/matcher/loftr_module/fine_preprocess.py:50-57);
    `w`: down_proj.weight/bias, merge_feat.weight/bias."""
    m, ww, _ = win0.shape
    c_win = F.linear(torch.cat([feat_c0[b_ids, i_ids], feat_c1[b_ids, j_ids]], 0).float(), w["down_proj.weight"],
                     w["down_proj.bias"])
    both = torch.cat([win0, win1], 0).float()
    merged = F.linear(torch.cat([both, c_win[:, None, :].expand(-1, ww, -1)], -1), w["merge_feat.weight"],
                      w["merge_feat.bias"])
    return merged[:m], merged[m:]


def cosine_scores(q: torch.Tensor, refs: torch.Tensor) -> torch.Tensor:
    """score[r] = F.cosine_similarity(q, refs[r:r+1], dim=1, eps=1e-8), one crop at a time, exactly as the
    retrieval loop does (eval_linemod_json.py:94; token from dinov2_utils.py:106-111)."""
    return torch.stack([F.cosine_similarity(q.float(), refs[r:r + 1].float(), dim=1, eps=1e-8)[0]
                        for r in range(refs.shape[0])])


def running_topk(scores: Sequence[float], k: int = 3) -> Tuple[list, list]:
    """The slot-replacement top-k of eval_linemod_json.py:72-101: k slots start at score 0; a crop whose
    score exceeds *any* slot overwrites the current arg-min slot (first arg-min on ties, numpy).
    Returns (slot_scores, slot_indices) with index -1 for a slot never filled."""
    slot_s = [0.0] * k
    slot_i = [-1] * k
    for r, sc in enumerate(scores):
        sc = float(sc)
        if any(sc > v for v in slot_s):
            lo = min(range(k), key=lambda t: (slot_s[t], t))
            slot_s[lo] = sc
            slot_i[lo] = r
    return slot_s, slot_i


def match_scores(mconf_per_pair: Sequence[torch.Tensor], group: int = 3, thr: float = 0.9) -> Tuple[list, list]:
    """eval_linemod_json.py:118-119 (`np.where(confidences > 0.9)[0].shape[0]` per crop) and :146 (`np.argmax` of the
    scores of one query's crops).  Returns (scores per pair, first arg-max per group of `group` consecutive pairs)."""
    scores = [int((np.asarray(c) > thr).sum()) for c in mconf_per_pair]
    best = [int(np.argmax(scores[g:g + group])) for g in range(0, len(scores), group)]
    return scores, best


def match_pairs(feat_c0, feat_c1, feat_f0, feat_f1, hw0_i, hw0_c, hw1_c, thr=THR, border_rm=BORDER_RM,
                temperature=DSMAX_TEMPERATURE, W=FINE_WINDOW) -> Dict[str, torch.Tensor]:
    """The hot path end to end without the fine transformer (SURVEY.md section 8(d) "hot path only"):
    coarse match -> window gather on both fine maps -> fine match on the raw windows."""
    out = coarse_match(feat_c0, feat_c1, hw0_i, hw0_c, hw1_c, thr, border_rm, temperature)
    stride = feat_f0.shape[2] // hw0_c[0]
    w0 = fine_windows(feat_f0, out["b_ids"], out["i_ids"], W, stride)
    w1 = fine_windows(feat_f1, out["b_ids"], out["j_ids"], W, stride)
    out.update(fine_match(w0, w1, out["mkpts0_c"], out["mkpts1_c"], hw0_i[0] / feat_f0.shape[2]))
    return out
