"""Mirror of reference core/utils/anchor.py (anchor / prior generators used by the decode path).

The fused kernels derive YOLOv8 anchor points from the cell index on the fly (make_anchors is never
materialised on the hot path); these host versions exist for callers that want the tables."""
from __future__ import annotations

import numpy as np
import torch


def make_anchors(feats, strides, grid_cell_offset=0.5):
    """YOLOv8 anchor points and stride column (reference :126-145): per level, cell centres
    (x + 0.5, y + 0.5) row-major, and the level stride repeated h*w times."""
    assert feats is not None
    dtype, device = feats[0].dtype, feats[0].device
    points, stride_col = [], []
    for level, stride in zip(feats, strides):
        h, w = level.shape[2], level.shape[3]
        xs = torch.arange(w, device=device, dtype=dtype) + grid_cell_offset
        ys = torch.arange(h, device=device, dtype=dtype) + grid_cell_offset
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        points.append(torch.stack((gx, gy), -1).reshape(-1, 2))
        stride_col.append(torch.full((h * w, 1), float(stride), dtype=dtype, device=device))
    return torch.cat(points), torch.cat(stride_col)


def generate_ssd_anchor_v2(input_image_shape, anchor_sizes, feature_shapes, aspect_ratios):
    """SSD prior boxes (reference :45-99, identical to Ssd._get_ssd_anchors ssd.py:482-541): computed in
    float64 like the reference and cast to float32; rows are (xmin, ymin, xmax, ymax) in [0, 1]."""
    image_h, image_w = input_image_shape
    out = []
    for i, fs in enumerate(feature_shapes):
        lo, hi = anchor_sizes[i], anchor_sizes[i + 1]
        ws, hs = [], []
        for ar in aspect_ratios[i]:
            if ar == 1:
                ws += [lo, np.sqrt(lo * hi)]
                hs += [lo, np.sqrt(lo * hi)]
            else:
                ws.append(lo * np.sqrt(ar))
                hs.append(lo / np.sqrt(ar))
        half_w, half_h = np.array(ws) / 2.0, np.array(hs) / 2.0
        step_y, step_x = image_h / fs, image_w / fs
        cx = np.linspace(0.5 * step_x, image_w - 0.5 * step_x, fs)
        cy = np.linspace(0.5 * step_y, image_h - 0.5 * step_y, fs)
        gx, gy = np.meshgrid(cx, cy)
        centres = np.concatenate((gx.reshape(-1, 1), gy.reshape(-1, 1)), axis=1)
        k = len(ws)
        boxes = np.tile(centres, (1, 2 * k))
        boxes[:, 0::4] -= half_w
        boxes[:, 1::4] -= half_h
        boxes[:, 2::4] += half_w
        boxes[:, 3::4] += half_h
        boxes[:, 0::2] /= image_w
        boxes[:, 1::2] /= image_h
        out.append(np.clip(boxes, 0.0, 1.0).reshape(-1, 4))
    return np.concatenate(out, axis=0).astype(np.float32)


def generate_yolo3_anchor(cfg, device, idx=None):
    """YOLOv3 anchors normalised by the input width/height (reference :102-117); rows 3*idx..3*idx+2
    belong to scale `idx`."""
    _, h, w = cfg.arch.input_size
    a = torch.tensor(cfg.arch.anchor, dtype=torch.float32).reshape(-1, 2)
    a[:, 0] /= w
    a[:, 1] /= h
    if device is not None:
        a = a.to(device)
    return a if idx is None else a[3 * idx: 3 * (idx + 1), :]


def get_yolo7_anchors(cfg):
    """YOLOv7 anchor table in pixels, shape (9, 2) (reference :120-123)."""
    return np.array(cfg.arch.anchors, dtype=np.float32).reshape(-1, 2)
