"""Mirror of reference core/utils/nms.py: `diou_nms` (:9-31), `gather_op` (:34-51), `yolo3_nms`
(:54-84) and `yolo7_nms` (:87-146), routed to libcvpp kernels."""
from __future__ import annotations

import torch

from ... import ops


def diou_nms(boxes, scores, iou_threshold):
    """Greedy class-agnostic DIoU-NMS.  boxes (N, 4) xyxy, scores (N,) -> int64 tensor of kept indices,
    sorted by decreasing score (the reference returns a CPU LongTensor; so does this)."""
    if boxes.numel() == 0:
        return torch.LongTensor([])
    return ops.diou_nms(boxes.float(), scores.float(), float(iou_threshold)).cpu()


def gather_op(tensor, indice, device):
    """rows of `tensor` (M,) or (M, N) at `indice` (K,) -> float32 (K, N) (reference :34-51, without its
    Python loop per row)."""
    assert tensor.dim() == 1 or tensor.dim() == 2
    src = (tensor if tensor.dim() == 2 else tensor[:, None]).float()
    if indice.numel() == 0:
        return torch.zeros((0, src.shape[1]), dtype=torch.float32, device=device)
    ind = indice.reshape(1, -1)
    if ind.dtype not in (torch.int32, torch.int64):
        ind = ind.to(torch.int64)
    return ops.gather_feat(src.unsqueeze(0), ind.to(src.device))[0].to(device)


def yolo3_nms(num_classes, conf_threshold, iou_threshold, boxes, scores, device):
    """boxes (M, 4) xyxy, scores (M, num_classes) -> (boxes (K, 4), scores (K, 1), classes (K,) int32):
    per class, `scores >= conf` then torchvision-style NMS; class ascending, score descending (reference
    :54-84): one candidate key per (row, class) above the threshold, then the shared sort + NMS kernels."""
    boxes = boxes.float().contiguous()
    scores = scores.float().contiguous()
    M = int(boxes.shape[0])
    if M == 0:
        return (torch.zeros((0, 4), dtype=torch.float32, device=device), torch.zeros((0, 1), dtype=torch.float32, device=device),
                torch.zeros((0,), dtype=torch.int32, device=device))
    cand = ops.score_matrix_filter(boxes, scores, float(conf_threshold))
    det = ops.per_class_nms_device(cand, float(iou_threshold))
    n = int(det.count.item())
    return det.box[0, :n].to(device), det.score[0, :n, None].to(device), det.cls[0, :n].to(device=device, dtype=torch.int32)


def yolo7_nms(prediction, num_classes, input_shape, image_shape, letterbox_image, device, conf_thres=0.5, nms_thres=0.4):
    """Free-function twin of YOLOv7._nms (reference :87-146)."""
    from ..algorithms.yolo_v7 import _rows_after_nms, _rows_to_list
    cand = ops.yolov7_pred_filter(prediction.float(), num_classes, conf_thres)
    B = int(cand.key.shape[0])
    rows, count = _rows_after_nms(cand, nms_thres, input_shape, [tuple(image_shape)] * B, letterbox_image)
    return _rows_to_list(rows, count)
