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
    src = tensor if tensor.dim() == 2 else tensor[:, None]
    return src.index_select(0, indice.to(torch.int64)).to(dtype=torch.float32, device=device)
