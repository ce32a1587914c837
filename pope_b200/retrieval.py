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


def retrieve_topk_sharded(local_scores: torch.Tensor, n_crops: int, rank: int, world: int, k: int = 3, group=None,
                          topk_fn=None) -> Tuple[torch.Tensor, List[float], List[int]]:
    """Retrieval with the R crops sharded over the ranks of one box: rank r scores the contiguous
    block `driver.shard_range(n_crops, r, world)` of the crops (ViT forwards + `ops.cosine_topk` on its own GPU -- that is
    where the time goes) and passes its scores [hi - lo] here.  ONE small collective, an all-gather of the R scores in crop
    order, then every rank runs the running top-k over all of them.  The slots of the reference loop
    (eval_linemod_json.py:95-101: a crop that beats any slot overwrites the current arg-min slot) depend on the whole arrival
    order -- an element that is evicted later still decides which slot its successor lands in -- so per-shard top-k lists
    cannot be merged into the same slots; the full score vector (R floats) can.
    Returns (scores [R], slot_scores, slot_indices), identical on every rank and to the single-GPU `retrieve_topk`.
    `topk_fn(scores, k) -> (slot_scores, slot_idx)`: defaults to the CUDA kernel (`ops.running_topk`)."""
    import torch.distributed as dist
    from .driver import shard_range
    spans = [shard_range(n_crops, r, world) for r in range(world)]
    lo, hi = spans[rank]
    if local_scores.numel() != hi - lo:
        raise ValueError(f"rank {rank} owns crops [{lo}, {hi}) but passed {local_scores.numel()} scores")
    width = max(b - a for a, b in spans)
    send = torch.zeros(width, dtype=torch.float32, device=local_scores.device)
    send[:hi - lo] = local_scores.float()
    if world > 1:
        recv = torch.empty(world * width, dtype=torch.float32, device=local_scores.device)
        dist.all_gather_into_tensor(recv, send, group=group)
        scores = torch.cat([recv[r * width:r * width + (b - a)] for r, (a, b) in enumerate(spans)])
    else:
        scores = send[:n_crops]
    slot_s, slot_i = (topk_fn or ops.running_topk)(scores, k)
    return scores, [float(v) for v in slot_s], [int(v) for v in slot_i]
