"""Deterministic synthetic head outputs for tests and benchmarks.

Raw head logits per level, `(B, no, H_l, W_l)` fp32 NCHW, channel order
`box[0:4*reg_max] | cls[nc] | embedding[embed_dim] | state[state_classes]`
(reference layout: ultralytics/nn/modules/head.py:204-206, :232-235).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def level_shapes(imgsz, strides: Sequence[int]) -> List[Tuple[int, int]]:
    """Feature-map sizes (H_l, W_l) of an `imgsz` (int or (h, w)) input at each stride."""
    h, w = (imgsz, imgsz) if isinstance(imgsz, int) else imgsz
    return [(-(-h // s), -(-w // s)) for s in strides]


def head_outputs(batch: int, shapes: Sequence[Tuple[int, int]], nc: int, embed_dim: int = 0, state_classes: int = 0,
                 reg_max: int = 16, cls_mean: float = -4.0, cls_std: float = 2.0, box_std: float = 2.0,
                 seed: int = 0, device="cpu", blobs: int = 0) -> List[torch.Tensor]:
    """Seeded N(0, box_std²) box logits, N(cls_mean, cls_std²) class logits, N(0,1) extras.

    `blobs` > 0 adds that many Gaussian bumps to the class logits of every image so that
    neighbouring anchors fire together and NMS really suppresses (clustered variant).
    """
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(int(seed))
    no = 4 * reg_max + nc + embed_dim + state_classes
    out = []
    for li, (h, w) in enumerate(shapes):
        x = torch.randn((batch, no, h, w), generator=g, device=dev, dtype=torch.float32)
        x[:, : 4 * reg_max] *= box_std
        x[:, 4 * reg_max: 4 * reg_max + nc] *= cls_std
        x[:, 4 * reg_max: 4 * reg_max + nc] += cls_mean
        if blobs:
            cy = torch.rand((batch, blobs, 1, 1), generator=g, device=dev) * h
            cx = torch.rand((batch, blobs, 1, 1), generator=g, device=dev) * w
            sig = 0.03 * max(h, w) + 0.5
            yy = torch.arange(h, device=dev, dtype=torch.float32).view(1, 1, h, 1)
            xx = torch.arange(w, device=dev, dtype=torch.float32).view(1, 1, 1, w)
            bump = (6.0 * torch.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * sig * sig))).sum(1, keepdim=True)
            x[:, 4 * reg_max: 4 * reg_max + nc] += bump
            # pull the box logits of fired anchors toward a common shape so their boxes overlap
            x[:, : 4 * reg_max] *= 1.0 / (1.0 + bump)
        out.append(x)
    return out


def decoded_prediction(batch: int, anchors: int, nc: int, nm: int = 0, seed: int = 0, device="cpu",
                       img: float = 640.0, score_pow: float = 4.0, clustered: bool = False) -> torch.Tensor:
    """A decoded `(B, 4+nc+nm, A)` tensor in the format `ops.non_max_suppression` takes
    (xywh pixels, class probabilities, extras) for NMS-only tests."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(int(seed))
    y = torch.empty((batch, 4 + nc + nm, anchors), device=dev, dtype=torch.float32)
    if clustered:
        k = max(anchors // 40, 1)
        centers = torch.rand((batch, 2, k), generator=g, device=dev) * img
        pick = torch.randint(0, k, (batch, anchors), generator=g, device=dev)
        cxy = torch.gather(centers, 2, pick.unsqueeze(1).expand(-1, 2, -1))
        y[:, 0:2] = cxy + torch.randn((batch, 2, anchors), generator=g, device=dev) * 6.0
        y[:, 2:4] = 40.0 + torch.rand((batch, 2, anchors), generator=g, device=dev) * 40.0
    else:
        y[:, 0:2] = torch.rand((batch, 2, anchors), generator=g, device=dev) * img
        y[:, 2:4] = 4.0 + torch.rand((batch, 2, anchors), generator=g, device=dev) * (img / 8)
    y[:, 4:4 + nc] = torch.rand((batch, nc, anchors), generator=g, device=dev) ** score_pow
    if nm:
        y[:, 4 + nc:] = torch.randn((batch, nm, anchors), generator=g, device=dev)
    return y
