"""Mirror of the decode entry points of reference core/predict/yolov3_decode.py: `predict_bounding_bbox`
(:12-29) and `Decoder` (:32-66), routed to libcvpp kernels."""
from __future__ import annotations

import torch

from ... import ops
from ..utils.anchor import generate_yolo3_anchor


def predict_bounding_bbox(num_classes, feature_map, anchors, device, is_training=False):
    """feature_map (N, 3*(nc+5), H, W), anchors (3, 2) normalised ->
    (box_xy (N,H,W,3,2), box_wh (N,H,W,3,2), confidence (N,H,W,3,1), class_prob (N,H,W,3,nc)); with
    is_training the last two are (grid (H,W,1,2), feature_map (N,H,W,3,nc+5)) like the reference (used by
    the YOLOv3 loss, core/loss/yolov3_loss.py:72)."""
    N, C, H, W = feature_map.size()
    box_xy, box_wh, confidence, class_prob = ops.yolov3_predict_bbox(feature_map.float(), num_classes,
                                                                     anchors.reshape(-1, 2).tolist())
    if not is_training:
        return box_xy, box_wh, confidence, class_prob
    gy, gx = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=device),
                            torch.arange(W, dtype=torch.float32, device=device), indexing="ij")
    grid = torch.stack((gx, gy), dim=-1).reshape(H, W, 1, 2)
    fm = feature_map.permute(0, 2, 3, 1).reshape(-1, H, W, 3, num_classes + 5)
    return box_xy, box_wh, grid, fm


class Decoder:
    """`Decoder(cfg, conf_threshold, device)(outputs)` -> (boxes (M, 4) normalised xyxy, scores (M,),
    classes (M,) int32), class ascending then score descending.  Like the reference, a batch is flattened
    into one set before NMS (yolov3_decode.py:47-50) - in effect a batch-1 API; `decode_batch` keeps the
    images apart."""

    def __init__(self, cfg, conf_threshold, device):
        self.cfg = cfg
        self.device = device
        self.num_classes = cfg.arch.num_classes
        self.conf_threshold = conf_threshold
        self.iou_threshold = cfg.decode.iou_threshold
        self._anchors_px = torch.tensor(cfg.arch.anchor, dtype=torch.float32).reshape(-1, 2).tolist()
        self.input_hw = cfg.arch.input_size[1:]

    def _yolo_post_process(self, feature, scale_type):
        """One scale, dense: (boxes (N*H*W*3, 4), scores (N*H*W*3, nc)) (reference :40-51)."""
        xy, wh, conf, prob = predict_bounding_bbox(self.num_classes, feature,
                                                   generate_yolo3_anchor(self.cfg, self.device, scale_type), self.device)
        boxes = torch.cat((xy - wh / 2, xy + wh / 2), dim=-1).reshape(-1, 4)
        return boxes, (conf * prob).reshape(-1, self.num_classes)

    def _candidates(self, outputs, merge_batch):
        ls = ops.make_levels([o.float() for o in outputs])
        n = ls.n
        cap = None
        while True:
            cand = ops.yolov3_decode_filter(ls, self.num_classes, self._anchors_px[:3 * n], self.input_hw,
                                            self.conf_threshold, merge_batch=merge_batch, max_cand=cap)
            need = int(cand.count.max().item()) if cand.count.numel() else 0
            if need <= cand.max_cand:
                return cand
            cap = need

    def __call__(self, outputs):
        cand = self._candidates(outputs, merge_batch=True)
        if cand.key.shape[0] == 0:
            return (torch.zeros((0, 4), device=self.device), torch.zeros((0,), device=self.device),
                    torch.zeros((0,), dtype=torch.int32, device=self.device))
        det = ops.per_class_nms_device(cand, self.iou_threshold)
        n = int(det.count.item())
        return det.box[0, :n], det.score[0, :n], det.cls[0, :n]

    def decode_batch(self, outputs) -> ops.Detections:
        """Batched extension: per-image candidates and NMS, detections left on the device."""
        return ops.per_class_nms_device(self._candidates(outputs, merge_batch=False), self.iou_threshold)
