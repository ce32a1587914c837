"""pope_b200 -- B200-native (sm_100a) implementation of the Matcher hot path of karltan0328/POPE.

    from pope_b200 import Matcher, default_cfg       # drop-in for `from src.matcher import Matcher, default_cfg`

The coarse-match / fine-window-gather / fine-match stages run in the hand-written CUDA library
`pope_b200/libpope_b200.so` (C ABI: include/pope_b200.h).  There is no CPU, PyTorch or Triton fallback:
calling the hot path without the built library or without a CUDA device raises.
"""
from .config import default_cfg, make_default_cfg          # noqa: F401
from .matcher import Matcher                               # noqa: F401
from .coarse_matching import CoarseMatching                # noqa: F401
from .fine_preprocess import FinePreprocess                # noqa: F401
from .fine_matching import FineMatching                    # noqa: F401
from .retrieval import retrieve_topk, retrieve_topk_images  # noqa: F401
from .dino_vit import DinoViT                              # noqa: F401
from .pose import compute_pose_errors, estimate_pose, estimate_pose_batch, relative_pose_error_batch  # noqa: F401

__all__ = ["Matcher", "default_cfg", "make_default_cfg", "CoarseMatching", "FinePreprocess", "FineMatching",
           "retrieve_topk", "retrieve_topk_images", "DinoViT", "estimate_pose", "estimate_pose_batch", "compute_pose_errors",
           "relative_pose_error_batch"]
