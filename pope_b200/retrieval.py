"""Reference-crop retrieval: cosine similarity of DINOv2 CLS tokens + the eval loop's running top-k, on device.

Replaces the per-crop `F.cosine_similarity(ref_fea, fea, dim=1, eps=1e-8)` + `.item()` + slot update of
eval_linemod_json.py:72-101 (token: segment_anything/segment_anything/dinov2_utils.py:106-111) by one call over
all R crops.  The ViT forward that produces the tokens is outside the accelerated path."""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import ops


def retrieve_topk(ref_fea: torch.Tensor, crop_feas: torch.Tensor, k: int = 3, eps: float = 1e-8
                  ) -> Tuple[torch.Tensor, List[float], List[int]]:
    """ref_fea [1,D] (the prompt image's token), crop_feas [R,D] in crop order.
    Returns (scores [R] on device, slot_scores, slot_indices) -- the slots are what the eval loop leaves in
    `similarity_score` / `top_images` (index -1: slot never filled, score stays 0)."""
    scores, slot_s, slot_i = ops.cosine_topk(ref_fea, crop_feas, k, eps)
    return scores, slot_s.tolist(), slot_i.tolist()


def retrieve_topk_images(model, ref_image: torch.Tensor, crop_images: torch.Tensor, k: int = 3, batch: int = 128,
                         eps: float = 1e-8) -> Tuple[torch.Tensor, List[float], List[int]]:
    """The whole retrieval step of eval_linemod_json.py:65-101 for one query: ref_image [1,3,H,W] (the prompt), crop_images
    [R,3,H,W] (the SAM crops after `set_torch_image`), `model` a DINOv2 ViT (pope_b200.dino_vit.DinoViT or the reference's
    own module).  ceil(R / batch) + 1 forwards instead of R + 1 batch-1 forwards, no per-crop host synchronisation; the
    slot semantics of the running top-k are those of the loop (crop order matters)."""
    from .dino_vit import cls_tokens
    ref_fea = cls_tokens(model, ref_image, batch)
    crop_feas = cls_tokens(model, crop_images, batch)
    return retrieve_topk(ref_fea, crop_feas, k, eps)
