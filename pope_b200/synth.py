"""Seeded synthetic inputs at the stage boundary of the Matcher hot path.

Images are useless for exercising the path with random-init weights (the reference Matcher yields
M = 0 matches on any image pair, SURVEY.md section 7 "vacuous parity trap"), so parity tests, `smoke()`
and `bench.py` drive the path with synthetic *features* carrying planted correspondences
(SURVEY.md section 8(d)):  f0 = sigma*randn, f1 = sigma*randn with a fraction of f1's rows overwritten by
noisy copies of f0's rows.  sigma = 1 gives M ~ 0.53*L matches with mconf spread over [0.2, 0.99].

Everything here is plain torch on the CPU generator so the same seed gives the same tensors on the
build container, the GPU box and inside the golden-vector script.
"""
from __future__ import annotations

import torch


def coarse_features(seed: int, n_pairs: int, L: int, S: int, C: int = 256, sigma: float = 1.0,
                    planted: float = 0.7, noise: float = 0.3, dtype=torch.float32):
    """Returns (feat_c0 [N,L,C], feat_c1 [N,S,C]) on the CPU.

    With dtype=torch.bfloat16 the values are rounded once to bf16; callers that need the fp32 oracle on
    the *same rounded values* use `.float()` on the result (SURVEY.md section 7 "bf16 parity protocol").
    """
    g = torch.Generator().manual_seed(seed)
    f0 = sigma * torch.randn(n_pairs, L, C, generator=g)
    f1 = sigma * torch.randn(n_pairs, S, C, generator=g)
    k = int(planted * min(L, S))
    for n in range(n_pairs):
        dst = torch.randperm(S, generator=g)[:k]
        src = torch.randperm(L, generator=g)[:k]
        f1[n, dst] = f0[n, src] + noise * torch.randn(k, C, generator=g)
    return f0.to(dtype), f1.to(dtype)


def hard_coarse_features(seed: int, n_pairs: int, L: int, S: int, C: int = 256, sigma: float = 1.0,
                         dtype=torch.float32):
    """The 'hard set': duplicated reference rows (two candidates per query row / exact ties), near
    duplicates, and one row per side scaled x50 for dynamic range (|S| reaches several hundred)."""
    f0, f1 = coarse_features(seed, n_pairs, L, S, C, sigma=sigma)
    g = torch.Generator().manual_seed(seed + 7919)
    for n in range(n_pairs):
        a = torch.randperm(S, generator=g)[: max(2, S // 40)]
        b = torch.randperm(S, generator=g)[: a.numel()]
        half = a.numel() // 2
        f1[n, a[:half]] = f1[n, b[:half]]                                  # exact duplicates
        f1[n, a[half:]] = f1[n, b[half:]] + 1e-3 * torch.randn(a.numel() - half, C, generator=g)
        big0 = torch.randperm(L, generator=g)[:1]
        big1 = torch.randperm(S, generator=g)[:1]
        f0[n, big0] *= 50.0
        f1[n, big1] *= 50.0
    return f0.to(dtype), f1.to(dtype)


def fine_feature_maps(seed: int, n_pairs: int, hf: int, wf: int, C: int = 128, dtype=torch.float32,
                      channels_last: bool = True):
    """Two fine-level feature maps, logical shape [N, C, hf, wf] (the backbone's 1/2-resolution output).

    channels_last=True returns torch.channels_last strides (C contiguous), the layout the CUDA window
    gather is fastest on; False gives plain NCHW like the reference backbone emits."""
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(n_pairs, C, hf, wf, generator=g).to(dtype)
    b = torch.randn(n_pairs, C, hf, wf, generator=g).to(dtype)
    if channels_last:
        a = a.contiguous(memory_format=torch.channels_last)
        b = b.contiguous(memory_format=torch.channels_last)
    return a, b


def fine_windows(seed: int, M: int, WW: int = 25, C: int = 128, dtype=torch.float32, gain: float = 1.0):
    """Post-transformer windows for the FineMatching-only tests: two [M, WW, C] tensors."""
    g = torch.Generator().manual_seed(seed)
    return (gain * torch.randn(M, WW, C, generator=g)).to(dtype), \
           (gain * torch.randn(M, WW, C, generator=g)).to(dtype)


def retrieval_tokens(seed: int, R: int = 256, D: int = 384, dtype=torch.float32):
    """One query CLS token [1, D] and R reference-crop CLS tokens [R, D]; a few references are noisy
    copies of the query so the top-k is not a coin flip."""
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(1, D, generator=g)
    refs = torch.randn(R, D, generator=g)
    hot = torch.randperm(R, generator=g)[: max(3, R // 16)]
    refs[hot] = q + torch.linspace(0.4, 2.0, hot.numel()).unsqueeze(1) * torch.randn(hot.numel(), D, generator=g)
    return q.to(dtype), refs.to(dtype)


def posed_pair_features(seed: int, n_pairs: int, hw_c=(60, 80), C: int = 256, Cf: int = 128, noise: float = 0.3,
                        focal: float = 500.0, max_angle: float = 0.12, dtype=torch.float32):
    """Feature pairs whose planted correspondences obey a two-view geometry, for the full chain match -> pose.

    Every coarse cell of image 0 (pixel (8x, 8y), the reference's mkpts0_c) gets a random depth and is moved by a random
    rigid motion (R, t) per pair; where its projection lands inside image 1, the coarse feature of the nearest cell of
    image 1 becomes a noisy copy (as in coarse_features) and the fine map of image 1 receives the centre vector of the
    image-0 window at the fine pixel (2 px) nearest to the projection, so FineMatching's expectation lands within about a
    pixel of the true projection.  Returns dict(feat_c0, feat_c1, feat_f0, feat_f1 (channels-last), K [3,3], R [n,3,3],
    t [n,3] (unit), proj [n, L, 2] (true projections, NaN where none was planted))."""
    import math
    g = torch.Generator().manual_seed(seed)
    hc, wc = hw_c
    L = hc * wc
    hf, wf = hc * 4, wc * 4
    K = torch.tensor([[focal, 0.0, wc * 4.0], [0.0, focal, hc * 4.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
    f0 = torch.randn(n_pairs, L, C, generator=g)
    f1 = torch.randn(n_pairs, L, C, generator=g)
    ff0 = torch.randn(n_pairs, hf, wf, Cf, generator=g)         # NHWC storage
    ff1 = torch.randn(n_pairs, hf, wf, Cf, generator=g)
    ys, xs = torch.meshgrid(torch.arange(hc), torch.arange(wc), indexing="ij")
    pix = torch.stack([xs.reshape(-1) * 8.0, ys.reshape(-1) * 8.0, torch.ones(L)], 1).to(torch.float64)     # [L, 3]
    rays = pix @ torch.linalg.inv(K).T
    Rs, ts, projs = [], [], []
    for n in range(n_pairs):
        ax = torch.randn(3, generator=g, dtype=torch.float64)
        ax = ax / ax.norm()
        ang = float(torch.rand(1, generator=g)) * max_angle + 0.03
        kx = torch.tensor([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]], dtype=torch.float64)
        R = torch.eye(3, dtype=torch.float64) + math.sin(ang) * kx + (1 - math.cos(ang)) * (kx @ kx)
        t = torch.randn(3, generator=g, dtype=torch.float64)
        t = t / t.norm()
        z = 3.0 + 3.0 * torch.rand(L, generator=g, dtype=torch.float64)
        X1 = (rays * z[:, None]) @ R.T + 0.4 * t
        uv = (X1 @ K.T)
        uv = uv[:, :2] / uv[:, 2:]
        jx, jy = torch.round(uv[:, 0] / 8.0).long(), torch.round(uv[:, 1] / 8.0).long()
        ok = (X1[:, 2] > 0) & (jx >= 0) & (jx < wc) & (jy >= 0) & (jy < hc)
        fx, fy = torch.round(uv[:, 0] / 2.0).long(), torch.round(uv[:, 1] / 2.0).long()          # nearest fine pixel
        ok &= (fx >= 0) & (fx < wf) & (fy >= 0) & (fy < hf) & ((fx - 4 * jx).abs() <= 2) & ((fy - 4 * jy).abs() <= 2)
        j = jy * wc + jx
        src = torch.nonzero(ok).reshape(-1)
        seen = torch.zeros(L, dtype=torch.bool)
        keep = []
        for i in src.tolist():              # first source cell wins a target cell
            if not seen[j[i]]:
                seen[j[i]] = True
                keep.append(i)
        keep = torch.tensor(keep, dtype=torch.long)
        f1[n, j[keep]] = f0[n, keep] + noise * torch.randn(keep.numel(), C, generator=g)
        cy, cx = (keep // wc) * 4, (keep % wc) * 4                                               # window centres in map 0
        ff1[n, fy[keep], fx[keep]] = ff0[n, cy, cx]
        proj = torch.full((L, 2), float("nan"), dtype=torch.float64)
        proj[keep] = uv[keep]
        Rs.append(R); ts.append(t); projs.append(proj)
    return dict(feat_c0=f0.to(dtype), feat_c1=f1.to(dtype), feat_f0=ff0.to(dtype).permute(0, 3, 1, 2),
                feat_f1=ff1.to(dtype).permute(0, 3, 1, 2), K=K, R=torch.stack(Rs), t=torch.stack(ts), proj=torch.stack(projs))
