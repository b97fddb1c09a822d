"""Mirror of the decode entry points of reference core/algorithms/yolo_v7.py: `YOLOv7.decode_box`
(:234-346), `YOLOv7._nms` (:348-422), `YOLOv7.get_anchors` (:45-49), plus a batched device-side
extension (`decode_batch`) with one (h, w) per image."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from ... import ops


def _rows_after_nms(cand: ops.Candidates, nms_thres: float, input_hw, image_hw: Sequence[Tuple[int, int]],
                    letterbox_image: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """sort -> per-class NMS -> (B, max_out, 7) rows [x1,y1,x2,y2,obj,class_conf,cls] in original-image pixels
    (yolo_v7.py:391-421), still on the device, plus the per-image counts."""
    det = ops.per_class_nms_device(cand, nms_thres)
    table = ops.correct_boxes_params(image_hw, input_hw, letterbox_image, det.box.device)
    rows = ops.detection_epilogue(det, ops.ROWS_YOLOV7, ops.BOX_CORRECT, table, cand.aux_dense)
    return rows, det.count


def _rows_to_list(rows: torch.Tensor, count: torch.Tensor) -> List[Optional[np.ndarray]]:
    """One device->host transfer; an image without detections yields None like the reference."""
    rows_h, counts = rows.cpu().numpy(), count.cpu().tolist()
    return [rows_h[b, :n].copy() if n > 0 else None for b, n in enumerate(counts)]


class YOLOv7:
    """Decode half of the reference's YOLOv7 algorithm class.  `cfg` is the reference's Yolo7Config (or any
    object with the same attributes)."""

    def __init__(self, cfg, device):
        self.cfg = cfg
        self.device = device
        self.anchors = self.get_anchors()
        self.num_classes = cfg.dataset.num_classes
        self.input_image_size = cfg.arch.input_size[1:]
        self.bbox_attrs = 5 + self.num_classes
        self.anchors_mask = cfg.arch.anchors_mask
        self.letterbox_image = cfg.decode.letterbox_image
        self.conf_threshold = cfg.decode.conf_threshold
        self.nms_threshold = cfg.decode.nms_threshold

    def get_anchors(self) -> np.ndarray:
        return np.array(self.cfg.arch.anchors, dtype=np.float32).reshape(-1, 2)

    def _level_anchors(self, n_levels: int) -> np.ndarray:
        return np.concatenate([self.anchors[list(self.anchors_mask[i])] for i in range(n_levels)], axis=0)

    def _candidates(self, preds, conf_threshold) -> ops.Candidates:
        levels = [p.float() for p in preds]
        ls = ops.make_levels(levels)
        return ops.yolov7_decode_filter(ls, self.num_classes, self._level_anchors(ls.n), self.input_image_size,
                                        conf_threshold)

    def decode_box(self, preds, image_h, image_w, conf_threshold=None):
        """preds: the three head levels (B, 3*(5+nc), 20|40|80, ...) -> list of B float32 ndarrays (n_i, 7)
        [x1, y1, x2, y2, obj, class_conf, class_id] in original-image pixels (None when an image has no
        detection), class-ascending then score-descending.  Every image gets the same (image_h, image_w)
        like the reference."""
        if conf_threshold is None:
            conf_threshold = self.conf_threshold
        cand = self._candidates(preds, conf_threshold)
        B = int(cand.key.shape[0])
        rows, count = _rows_after_nms(cand, self.nms_threshold, self.input_image_size, [(image_h, image_w)] * B,
                                      self.letterbox_image)
        return _rows_to_list(rows, count)

    def decode_batch(self, preds, image_hw: Sequence[Tuple[int, int]], conf_threshold=None):
        """Batched extension: per-image original sizes; returns (rows (B, max_out, 7), count (B,)) on the device."""
        if conf_threshold is None:
            conf_threshold = self.conf_threshold
        cand = self._candidates(preds, conf_threshold)
        return _rows_after_nms(cand, self.nms_threshold, self.input_image_size, image_hw, self.letterbox_image)

    def _nms(self, prediction, input_shape, image_shape, conf_threshold):
        """prediction (B, A, 5+nc) decoded (cx, cy, w, h, obj, cls...) normalised -> same list as decode_box."""
        cand = ops.yolov7_pred_filter(prediction.float(), self.num_classes, conf_threshold)
        B = int(cand.key.shape[0])
        rows, count = _rows_after_nms(cand, self.nms_threshold, input_shape, [tuple(image_shape)] * B,
                                      self.letterbox_image)
        return _rows_to_list(rows, count)
