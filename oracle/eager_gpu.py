"""TEST / BENCH INFRASTRUCTURE ONLY - never imported by the product (computervision/pytorch_b200).

The "reference on the GPU" bar: what the reference's own YOLOv8 path costs
when its tensors live on the B200 - eager ATen ops + torchvision's sm_100 `nms_kernel_impl` through
torchvision.ops.batched_nms.  The reference tree cannot travel to the GPU box, so its two functions are
restated here op for op in eager torch (same op order, one kernel per op, the same torchvision call):

  detect_tail           Detect.forward eval tail  core/models/yolov8/modules.py:434-445 (+ DFL :80-82,
                        make_anchors core/utils/anchor.py:126-145, dist2bbox core/utils/bboxes.py:213-222)
  non_max_suppression   core/utils/ultralytics_ops.py:131-264 (best-class branch, no time limit), with an
                        extra index channel so the kept ANCHOR indices come back

tests/test_eager_ref.py pins both against the fixtures the real reference wrote (tests/golden/yolov8_*.npz) on
CPU tensors, so what bench.py times on CUDA tensors is the reference's op sequence, not an approximation of it.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torchvision


def detect_tail(levels: Sequence[torch.Tensor], strides: Sequence[float], nc: int, reg_max: int = 16) -> torch.Tensor:
    """levels: list of (B, 4*reg_max + nc, H, W) -> y (B, 4 + nc, A)."""
    B = levels[0].shape[0]
    no = 4 * reg_max + nc
    dev, dt = levels[0].device, levels[0].dtype
    pts, sts = [], []
    for x, s in zip(levels, strides):                                    # make_anchors, offset 0.5
        h, w = x.shape[2], x.shape[3]
        sx = torch.arange(w, device=dev, dtype=dt) + 0.5
        sy = torch.arange(h, device=dev, dtype=dt) + 0.5
        gy, gx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((gx, gy), -1).view(-1, 2))
        sts.append(torch.full((h * w, 1), float(s), dtype=dt, device=dev))
    anchors = torch.cat(pts).transpose(0, 1)                             # (2, A)
    stride_t = torch.cat(sts).transpose(0, 1)                            # (1, A)
    x_cat = torch.cat([x.view(B, no, -1) for x in levels], 2)
    box, cls = x_cat.split((reg_max * 4, nc), 1)
    a = box.shape[2]
    # DFL: softmax over the bins of the transposed view, then the frozen 1x1 conv with weights arange(reg_max)
    w = torch.arange(reg_max, dtype=dt, device=dev).view(1, reg_max, 1, 1)
    dist = torch.nn.functional.conv2d(box.view(B, 4, reg_max, a).transpose(2, 1).softmax(1), w).view(B, 4, a)
    lt, rb = dist.chunk(2, 1)                                            # dist2bbox(xywh=True, dim=1)
    x1y1 = anchors.unsqueeze(0) - lt
    x2y2 = anchors.unsqueeze(0) + rb
    dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * stride_t
    return torch.cat((dbox, cls.sigmoid()), 1)


def _xywh2xyxy(x: torch.Tensor) -> torch.Tensor:
    y = x.clone()                                # ultralytics_ops.py:360-375: a copy, then four column updates,
    y[..., 0] = x[..., 0] - x[..., 2] / 2        # each with its own half-extent division
    y[..., 1] = x[..., 1] - x[..., 3] / 2
    y[..., 2] = x[..., 0] + x[..., 2] / 2
    y[..., 3] = x[..., 1] + x[..., 3] / 2
    return y


def non_max_suppression(pred: torch.Tensor, conf_thres: float, iou_thres: float, max_det: int = 300, nc: int = 0,
                        max_nms: int = 30000) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    """pred (B, 4 + nc, A) -> (rows per image (n, 6) = x1,y1,x2,y2,conf,cls; kept anchor indices per image (n,) int64).
    The branch torchvision.ops.batched_nms takes depends on the DEVICE of pred (boxes.py:80), exactly as it would for
    the reference."""
    B, ch, A = pred.shape
    nc = nc or ch - 4
    idx = torch.arange(A, device=pred.device, dtype=pred.dtype)[None, None].expand(B, 1, A)
    pred = torch.cat((pred[:, :4 + nc], idx), 1)                          # the index rides along as one extra channel
    xc = pred[:, 4:4 + nc].amax(1) > conf_thres
    rows: List[torch.Tensor] = []
    anchors: List[torch.Tensor] = []
    for xi in range(B):
        x = pred[xi].transpose(0, -1)[xc[xi]]
        if not x.shape[0]:
            rows.append(pred.new_zeros((0, 6)))
            anchors.append(torch.zeros((0,), dtype=torch.int64, device=pred.device))
            continue
        box, cls, extra = x.split((4, nc, 1), 1)
        box = _xywh2xyxy(box)
        conf, j = cls.max(1, keepdim=True)
        x = torch.cat((box, conf, j.float(), extra), 1)[conf.view(-1) > conf_thres]
        if not x.shape[0]:
            rows.append(pred.new_zeros((0, 6)))
            anchors.append(torch.zeros((0,), dtype=torch.int64, device=pred.device))
            continue
        x = x[x[:, 4].argsort(descending=True)[:max_nms]]
        keep = torchvision.ops.batched_nms(x[:, :4], x[:, 4], x[:, 5], iou_thres)[:max_det]
        x = x[keep]
        rows.append(x[:, :6])
        anchors.append(x[:, 6].to(torch.int64))
    return rows, anchors
