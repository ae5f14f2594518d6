"""Detect / JDE head decode behind the reference's own method signature.

`detect_inference(self, x)` and `jde_inference(self, x)` have the signature of
`Detect._inference` (ultralytics/nn/modules/head.py:100) and `JDE._inference` (:214) and can be
bound onto the reference classes (see plugin.patch).  `make_anchors` / `dist2bbox` mirror
ultralytics/utils/tal.py:366-390 for callers that use them directly; inside the fused kernels the
anchors are analytic (x = i % W + 0.5, y = i // W + 0.5) and never materialised.
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import ops as _ops


def _spec_of(self) -> _ops.HeadSpec:
    return _ops.HeadSpec.from_module(self)


def _check_export(self):
    if getattr(self, "export", False):
        raise NotImplementedError("sarpost: export-format decode branches (head.py:109-127) are not accelerated; "
                                  "unpatch before exporting")


def detect_inference(self, x: Sequence[torch.Tensor]) -> torch.Tensor:
    """Drop-in for `Detect._inference(self, x)` (head.py:100-131): returns `cat(dbox, cls.sigmoid())`."""
    _check_export(self)
    self.shape = x[0].shape  # head.py:107 keeps the anchor cache key; nothing else is cached here
    return _ops.decode(x, _spec_of(self))


def jde_inference(self, x: Sequence[torch.Tensor]) -> torch.Tensor:
    """Drop-in for `JDE._inference(self, x)` (head.py:214-249): returns
    `cat(dbox, cls.sigmoid(), emb, state.sigmoid())`."""
    _check_export(self)
    self.shape = x[0].shape
    return _ops.decode(x, _spec_of(self))


def make_anchors(feats, strides, grid_cell_offset=0.5):
    """utils/tal.py:366-378 — anchor points (A, 2) and stride tensor (A, 1) on the device of `feats`."""
    pts, st = [], []
    dtype, device = feats[0].dtype, feats[0].device
    for i, stride in enumerate(strides):
        h, w = feats[i].shape[2:] if isinstance(feats, (list, tuple)) else (int(feats[i][0]), int(feats[i][1]))
        sx = torch.arange(end=w, device=device, dtype=dtype) + grid_cell_offset
        sy = torch.arange(end=h, device=device, dtype=dtype) + grid_cell_offset
        sy, sx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((sx, sy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(stride), dtype=dtype, device=device))
    return torch.cat(pts), torch.cat(st)
