"""DINOv2 ViT (the `vit_small` / patch-14 configuration POPE retrieves crops with), restated for BATCHED inference.

SURVEY.md section 8(f) rank 2: the reference calls the ViT once per crop at batch 1 (`eval_linemod_json.py:74-93`,
`get_cls_token_torch`, segment_anything/segment_anything/dinov2_utils.py:106-111) and synchronises on every score; here
all R crops go through one forward and the tokens go straight to `pope_cosine_topk` on the device.

This is a feature producer OUTSIDE the accelerated hot path (stock PyTorch: cuDNN conv, cuBLAS Linears, SDPA), like the
ResNet-FPN in feature_net.py.  It reproduces the function of `DinoVisionTransformer.forward_features`
(dinov2/dinov2/models/vision_transformer.py:165-245, blocks: dinov2/dinov2/layers/block.py NestedTensorBlock in eval mode =
x + ls1(attn(norm1 x)); x + ls2(mlp(norm2 x)), attention.py:49-66, mlp.py, layer_scale.py, patch_embed.py) with the SAME
parameter names, so the reference's checkpoint loads unchanged (`tests/test_dino_vit.py` checks the state-dict keys and the
outputs against the unmodified reference).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Attention(nn.Module):
    def __init__(self, dim: int, heads: int):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim, bias=True)

    def forward(self, x):
        b, n, c = x.shape
        q, k, v = self.qkv(x).reshape(b, n, 3, self.num_heads, c // self.num_heads).permute(2, 0, 3, 1, 4)
        x = F.scaled_dot_product_attention(q, k, v)            # softmax(q k^T / sqrt(d)) v  (attention.py:56-62)
        return self.proj(x.transpose(1, 2).reshape(b, n, c))


class _Mlp(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))


class _LayerScale(nn.Module):
    def __init__(self, dim: int, init: float):
        super().__init__()
        self.gamma = nn.Parameter(init * torch.ones(dim))

    def forward(self, x):
        return x * self.gamma


class _Block(nn.Module):
    def __init__(self, dim: int, heads: int, mlp_ratio: float, init_values: float):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim, heads)
        self.ls1 = _LayerScale(dim, init_values)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))
        self.ls2 = _LayerScale(dim, init_values)

    def forward(self, x):
        x = x + self.ls1(self.attn(self.norm1(x)))
        return x + self.ls2(self.mlp(self.norm2(x)))


class _PatchEmbed(nn.Module):
    def __init__(self, patch: int, dim: int):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)          # [B, (H/p)(W/p), dim]


class DinoViT(nn.Module):
    """vit_small by default: embed 384, depth 12, 6 heads, patch 14, position table for 518 x 518 (37 x 37 patches)."""

    def __init__(self, img_size: int = 518, patch_size: int = 14, embed_dim: int = 384, depth: int = 12, num_heads: int = 6,
                 mlp_ratio: float = 4.0, init_values: float = 1.0):
        super().__init__()
        self.patch_size, self.embed_dim = patch_size, embed_dim
        n = (img_size // patch_size) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dim))
        self.mask_token = nn.Parameter(torch.zeros(1, embed_dim))       # unused at inference; part of the checkpoint
        self.patch_embed = _PatchEmbed(patch_size, embed_dim)
        self.blocks = nn.ModuleList([_Block(embed_dim, num_heads, mlp_ratio, init_values) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.normal_(self.cls_token, std=1e-6)

    def _pos(self, n_tokens: int, w: int, h: int, dtype) -> torch.Tensor:
        """Position table resampled to the input's patch grid (vision_transformer.py:165-189: bicubic on the square
        table with the +0.1 scale-factor offset of the original DINO code)."""
        n = self.pos_embed.shape[1] - 1
        if n_tokens - 1 == n and w == h:
            return self.pos_embed
        pe = self.pos_embed.float()
        side = int(math.sqrt(n))
        w0, h0 = w // self.patch_size + 0.1, h // self.patch_size + 0.1
        grid = F.interpolate(pe[:, 1:].reshape(1, side, side, -1).permute(0, 3, 1, 2), mode="bicubic",
                             scale_factor=(w0 / side, h0 / side))
        assert int(w0) == grid.shape[-2] and int(h0) == grid.shape[-1]
        return torch.cat([pe[:, :1], grid.permute(0, 2, 3, 1).reshape(1, -1, pe.shape[-1])], 1).to(dtype)

    def forward_features(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        _, _, w, h = x.shape
        t = self.patch_embed(x)
        t = torch.cat([self.cls_token.expand(t.shape[0], -1, -1).to(t.dtype), t], 1)
        t = t + self._pos(t.shape[1], w, h, t.dtype)
        for blk in self.blocks:
            t = blk(t)
        tn = self.norm(t)
        return {"x_norm_clstoken": tn[:, 0], "x_norm_patchtokens": tn[:, 1:], "x_prenorm": t, "masks": None}

    def forward(self, x: torch.Tensor, is_training: bool = False):
        """`model(x, is_training=True)` returns the feature dict, like the reference (vision_transformer.py:316-321);
        otherwise the CLS token."""
        out = self.forward_features(x)
        return out if is_training else out["x_norm_clstoken"]


@torch.no_grad()
def cls_tokens(model: nn.Module, images: torch.Tensor, batch: int = 128) -> torch.Tensor:
    """`get_cls_token_torch` (dinov2_utils.py:106-111) for a whole stack of crops [R, 3, H, W]: ceil(R / batch) forwards."""
    outs = [model(images[i:i + batch], is_training=True)["x_norm_clstoken"] for i in range(0, images.shape[0], batch)]
    return torch.cat(outs, 0)
