"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference Matcher from /root/reference.

This module exists so that `oracle/gen_golden.py` and `tests/test_oracle_vs_reference.py` can run the
reference's own code (src/matcher/...) in the build container, where /root/reference is mounted.  It is
never imported by the product package and nothing under `-m gpu`, `smoke()` or `bench.py` may use it
(/root/reference does not exist on the GPU box).

Two third-party packages the reference imports are absent from this image and cannot be installed
(no network): `yacs` (src/matcher/utils/cvpr_ds_config.py:1) and `kornia`
(src/matcher/utils/fine_matching.py:5-6).  We register minimal stand-ins in `sys.modules`:

* yacs.config.CfgNode -- the reference only assigns attributes, iterates `.items()` and uses
  `isinstance(., CN)` (cvpr_ds_config.py:4-50).
* kornia.utils.grid.create_meshgrid / kornia.geometry.subpix.dsnt.spatial_expectation2d -- restated from
  kornia's published semantics: a normalised meshgrid over [-1, 1] with (x, y) in the last dim, x varying
  fastest, and the expectation  sum(grid * heatmap)  over the flattened window.  kornia is not pinned by
  the reference (requirements.txt has no kornia line; upstream LoFTR pinned 0.4.1); for W=5 the
  semantics are identical across versions.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("POPE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "matcher"))


class _CfgNode(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=torch.float32):
    xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
    ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
    if normalized_coordinates:
        xs = (xs / (width - 1) - 0.5) * 2
        ys = (ys / (height - 1) - 0.5) * 2
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([gx, gy], dim=-1).unsqueeze(0)  # [1, H, W, 2], last dim (x, y)


def _spatial_expectation2d(inp, normalized_coordinates=True):
    b, n, h, w = inp.shape
    grid = _create_meshgrid(h, w, normalized_coordinates, inp.device).to(inp.dtype)
    flat = inp.reshape(b, n, -1)
    ex = torch.sum(grid[..., 0].reshape(-1) * flat, -1, keepdim=True)
    ey = torch.sum(grid[..., 1].reshape(-1) * flat, -1, keepdim=True)
    return torch.cat([ex, ey], -1)


def install_shims() -> None:
    if "yacs" not in sys.modules:
        yacs = types.ModuleType("yacs")
        cfg = types.ModuleType("yacs.config")
        cfg.CfgNode = _CfgNode
        yacs.config = cfg
        sys.modules["yacs"] = yacs
        sys.modules["yacs.config"] = cfg
    if "kornia" not in sys.modules:
        k = types.ModuleType("kornia")
        kg = types.ModuleType("kornia.geometry")
        ks = types.ModuleType("kornia.geometry.subpix")
        kd = types.ModuleType("kornia.geometry.subpix.dsnt")
        ku = types.ModuleType("kornia.utils")
        kug = types.ModuleType("kornia.utils.grid")
        kd.spatial_expectation2d = _spatial_expectation2d
        ks.dsnt = kd
        kg.subpix = ks
        kug.create_meshgrid = _create_meshgrid
        ku.grid = kug
        ku.create_meshgrid = _create_meshgrid
        k.geometry = kg
        k.utils = ku
        for name, mod in [("kornia", k), ("kornia.geometry", kg), ("kornia.geometry.subpix", ks),
                          ("kornia.geometry.subpix.dsnt", kd), ("kornia.utils", ku),
                          ("kornia.utils.grid", kug)]:
            sys.modules[name] = mod


def import_reference():
    """Returns the reference's `src.matcher` package (Matcher, default_cfg) imported unmodified."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    return importlib.import_module("src.matcher")
