"""Matcher configuration: a plain nested dict with the keys and values the reference resolves
`default_cfg` to (src/matcher/utils/cvpr_ds_config.py:10-50, lower-cased by `lower_config` :4-7).
No yacs dependency; `tests/test_host_logic.py` checks it against the reference's own dict (golden)."""
from __future__ import annotations

import copy


def make_default_cfg() -> dict:
    return {
        "backbone_type": "ResNetFPN",
        "resolution": (8, 2),                 # coarse 1/8, fine 1/2
        "fine_window_size": 5,
        "fine_concat_coarse_feat": True,
        "resnetfpn": {"initial_dim": 128, "block_dims": [128, 196, 256]},
        "coarse": {"d_model": 256, "d_ffn": 256, "nhead": 8, "layer_names": ["self", "cross"] * 4,
                   "attention": "linear", "temp_bug_fix": False},
        "match_coarse": {"thr": 0.2, "border_rm": 2, "match_type": "dual_softmax", "dsmax_temperature": 0.1,
                         "skh_iters": 3, "skh_init_bin_score": 1.0, "skh_prefilter": True,
                         "train_coarse_percent": 0.4, "train_pad_num_gt_min": 200},
        "fine": {"d_model": 128, "d_ffn": 128, "nhead": 8, "layer_names": ["self", "cross"], "attention": "linear"},
    }


default_cfg = make_default_cfg()


def clone_cfg(cfg: dict) -> dict:
    return copy.deepcopy(cfg)
