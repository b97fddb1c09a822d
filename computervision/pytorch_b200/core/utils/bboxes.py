"""Mirror of the box converters of reference core/utils/bboxes.py that the decode path uses
(`dist2bbox` :213-222, `xywh_to_xyxy_torch` :29-49, `xywh_to_xyxy` :9-26).  The fused kernels inline
this arithmetic; these are for callers holding plain tensors."""
from __future__ import annotations

import numpy as np
import torch


def xywh_to_xyxy(coords):
    """numpy (cx, cy, w, h) -> (xmin, ymin, xmax, ymax) on the last axis."""
    c, half = coords[..., 0:2], coords[..., 2:4] / 2
    return np.concatenate((c - half, c + half), axis=-1)


def xywh_to_xyxy_torch(coords, more=False):
    """torch (cx, cy, w, h, ...) -> (xmin, ymin, xmax, ymax, ...); `more` keeps the trailing columns."""
    c, half = coords[..., 0:2], coords[..., 2:4] / 2
    out = torch.cat((c - half, c + half), dim=-1)
    return torch.cat((out, coords[..., 4:]), dim=-1) if more else out


def dist2bbox(distance, anchor_points, xywh=True, dim=-1):
    """(l, t, r, b) distances from anchor points -> boxes (xywh or xyxy)."""
    lt, rb = distance.chunk(2, dim)
    tl, br = anchor_points - lt, anchor_points + rb
    if not xywh:
        return torch.cat((tl, br), dim)
    return torch.cat(((tl + br) / 2, br - tl), dim)
