"""Mirror of reference core/utils/ultralytics_ops.py for the detection path:
`non_max_suppression` (:131-264) and `xywh2xyxy` (:360-375).

`non_max_suppression` keeps the reference signature.  It runs two CUDA kernels through the C ABI
(confidence filter -> fused per-class sort + class-aware NMS) and reads the per-image counts back once;
per-image results are views of one row buffer (no per-image kernels) and there is no torchvision call.  The dead branches of the reference that no
caller reaches (multi_label, autolabelling `labels`, merge-NMS) raise NotImplementedError instead of
silently doing something else; `agnostic` is accepted and ignored exactly like the reference, whose
class-offset line is commented out (:243-247); `max_time_img` is accepted and ignored (the wall-clock
abort is a parity hazard)."""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from ... import ops


def xywh2xyxy(x):
    """(cx, cy, w, h) -> (x1, y1, x2, y2) on the last axis; torch.Tensor or np.ndarray (reference :360-375)."""
    y = x.clone() if isinstance(x, torch.Tensor) else np.copy(x)
    half_w, half_h = x[..., 2] / 2, x[..., 3] / 2
    y[..., 0] = x[..., 0] - half_w
    y[..., 1] = x[..., 1] - half_h
    y[..., 2] = x[..., 0] + half_w
    y[..., 3] = x[..., 1] + half_h
    return y


def _rows_from_detections(prediction: torch.Tensor, det: ops.Detections, nc: int, nm: int) -> List[torch.Tensor]:
    """(n_i, 6 + nm) rows [x1, y1, x2, y2, conf, cls, masks...] per image, like reference :226,257: ONE epilogue
    kernel assembles the (B, max_det, 6) rows on the device, one device->host read fetches the counts, and every
    image's result is a view of that buffer (mask columns, which no caller of the reference uses, are appended
    per image)."""
    B = prediction.shape[0]
    rows = ops.detection_epilogue(det, ops.ROWS_YOLOV8)
    counts = det.count.tolist()          # the one device->host read
    cap = det.box.shape[1]
    empty = torch.zeros((0, 6 + nm), device=prediction.device)
    out = [empty] * B                    # the reference aliases one empty tensor B times (:200)
    for b, n in enumerate(counts):
        if n <= 0:
            continue
        n = min(n, cap)
        out[b] = rows[b, :n] if not nm else \
            torch.cat((rows[b, :n], prediction[b, 4 + nc:, det.anchor[b, :n].long()].T), 1)
    return out


def non_max_suppression(
        prediction,
        conf_thres=0.25,
        iou_thres=0.45,
        classes=None,
        agnostic=False,
        multi_label=False,
        labels=(),
        max_det=300,
        nc=0,  # number of classes (optional)
        max_time_img=0.05,
        max_nms=30000,
        max_wh=7680,
        *,
        rule: int = ops.RULE_TORCHVISION_CPU,
        return_anchors: bool = False,
):
    """Confidence filter + class-aware NMS on a decoded prediction (B, 4 + nc + nm, A).

    Returns a list of B tensors (n_i, 6 + nm): x1, y1, x2, y2, confidence, class, masks..., sorted by
    confidence.  Keyword-only extensions: `rule` selects the torchvision.batched_nms arithmetic branch
    (default: the CPU rule the oracle follows), `return_anchors` additionally returns the kept anchor
    indices per image.
    """
    assert 0 <= conf_thres <= 1, f'Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0'
    assert 0 <= iou_thres <= 1, f'Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0'
    if isinstance(prediction, (list, tuple)):  # (inference_out, loss_out)
        prediction = prediction[0]
    if multi_label and (nc or prediction.shape[1] - 4) > 1:
        raise NotImplementedError("multi_label=True is a dead branch of the reference (no caller sets it)")
    if labels is not None and len(labels):
        raise NotImplementedError("autolabelling `labels` is a dead branch of the reference (no caller sets it)")
    if not prediction.is_cuda:
        raise ValueError("non_max_suppression runs on the GPU only: move `prediction` to a CUDA device")
    prediction = prediction.float()
    nc = nc or (prediction.shape[1] - 4)
    nm = prediction.shape[1] - nc - 4

    cand = ops.pred_filter(prediction, nc, conf_thres)
    if classes is not None:
        ops.keep_classes(cand, classes)      # reference :229-230
    det = ops.sort_nms(cand, iou_thres, rule, ops.ORDER_SCORE_DESC, max_det=max_det, max_nms=max_nms, max_out=max_det)
    rows = _rows_from_detections(prediction.contiguous(), det, nc, nm)
    if return_anchors:
        counts = det.count.tolist()
        return rows, [det.anchor[b, :n] for b, n in enumerate(counts)]
    return rows
