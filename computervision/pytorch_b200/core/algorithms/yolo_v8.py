"""Mirror of the decode entry point of reference core/algorithms/yolo_v8.py: `YOLOv8.decode_box`
(:210-242), plus two batched extensions the reference lacks (its decode asserts batch == 1)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from ... import ops
from ..utils.image_process import yolo_correct_boxes
from ..utils.ultralytics_ops import non_max_suppression


class YOLOv8:
    """Decode half of the reference's YOLOv8 algorithm class (model building / loss / training are out
    of scope).  `cfg` is the reference's Yolo8DetConfig (or any object with the same attributes)."""

    def __init__(self, cfg, device):
        self.cfg = cfg
        self.device = device
        self.num_classes = cfg.dataset.num_classes
        self.input_image_size = cfg.arch.input_size[1:]
        self.conf_threshold = cfg.decode.conf_threshold
        self.iou_threshold = cfg.decode.nms_threshold
        self.max_det = cfg.decode.max_det
        self.letterbox_image = cfg.decode.letterbox_image

    # -- reference signature ------------------------------------------------------------------
    def decode_box(self, preds, image_h, image_w, conf_threshold=None):
        """preds: decoded head output (1, 4 + nc, A) (or the (y, x) tuple) -> (bbox (n, 4) in pixels of
        the original image, conf (n,), cls (n,) int)."""
        if conf_threshold is None:
            conf_threshold = self.conf_threshold
        rows = non_max_suppression(preds, conf_threshold, self.iou_threshold, agnostic=False, max_det=self.max_det,
                                   classes=None)
        assert len(rows) == 1, "only a single image is supported by decode_box (use decode_batch)"
        return self._rows_to_image_space(rows[0].cpu().numpy(), image_h, image_w)

    # -- extensions ---------------------------------------------------------------------------
    def decode_batch(self, preds, image_hw: Sequence[Tuple[int, int]], conf_threshold=None):
        """Batched decode_box: one (bbox, conf, cls) triple per image, `image_hw[i] = (h, w)`."""
        if conf_threshold is None:
            conf_threshold = self.conf_threshold
        rows = non_max_suppression(preds, conf_threshold, self.iou_threshold, agnostic=False, max_det=self.max_det,
                                   classes=None)
        assert len(rows) == len(image_hw)
        return [self._rows_to_image_space(r.cpu().numpy(), h, w) for r, (h, w) in zip(rows, image_hw)]

    def decode_head(self, feats: Sequence[torch.Tensor], strides: Sequence[float], conf_threshold=None,
                    postprocessor: Optional[ops.Yolov8Postprocessor] = None) -> ops.Detections:
        """Fused path from the raw head levels (B, 4*16 + nc, H, W): decode + filter + sort + NMS in one
        C call, detections left on the device (input-image pixels, xyxy)."""
        if conf_threshold is None:
            conf_threshold = self.conf_threshold
        ls = ops.make_levels(list(feats), strides)
        post = postprocessor or ops.Yolov8Postprocessor(ls.B, ls.A, self.num_classes, ls.device, max_det=self.max_det)
        return post(ls, conf_threshold, self.iou_threshold)

    def _rows_to_image_space(self, pred: np.ndarray, image_h, image_w):
        bbox, conf, cls = pred[:, :4], pred[:, 4], pred[:, 5].astype(int)
        bbox[:, 0::2] /= self.input_image_size[1]
        bbox[:, 1::2] /= self.input_image_size[0]
        centre, size = (bbox[:, 0:2] + bbox[:, 2:4]) / 2, bbox[:, 2:4] - bbox[:, 0:2]
        bbox[:, :4] = yolo_correct_boxes(centre, size, self.input_image_size, [image_h, image_w], self.letterbox_image)
        return bbox, conf, cls
