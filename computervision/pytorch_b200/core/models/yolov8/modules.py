"""Mirror of the inference tail of reference core/models/yolov8/modules.py: `DFL` (:67-83) and the
eval branch of `Detect.forward` (:434-445).  The convolutions of the head are out of scope (cuDNN
territory); what is here is everything after them."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
from torch import nn

from .... import ops


class DFL(nn.Module):
    """Integral module of Distribution Focal Loss: softmax over the c1 bins, expectation with weights
    arange(c1).  (B, 4*c1, A) -> (B, 4, A).  Plain torch: the fused kernel inlines this on the hot path;
    the module is kept for the training loss, which calls it (reference core/algorithms/yolo_v8.py:73)."""

    def __init__(self, c1=16):
        super().__init__()
        self.c1 = c1
        self.register_buffer("bins", torch.arange(c1, dtype=torch.float32), persistent=False)

    def forward(self, x):
        b, _, a = x.shape
        p = x.view(b, 4, self.c1, a).softmax(2)
        return (p * self.bins.view(1, 1, self.c1, 1)).sum(2)


def detect_decode(x: Sequence[torch.Tensor], strides: Sequence[float], nc: int, reg_max: int = 16) -> torch.Tensor:
    """Eval tail of Detect.forward (:438-445) as ONE kernel: x = per-level (B, 4*reg_max + nc, H, W) ->
    y (B, 4 + nc, A) = cat(dist2bbox(DFL(box), anchors) * strides, sigmoid(cls))."""
    ls = ops.make_levels(list(x), strides)
    return ops.yolov8_decode_full(ls, nc, reg_max)


class Detect(nn.Module):
    """Inference tail of the YOLOv8 Detect head.  `forward(x)` takes the per-level tensors produced by
    the head convolutions (already concatenated box|cls channels, reference :431) and returns
    `(y, x)` like the reference's eval branch."""

    def __init__(self, nc=80, stride=(8.0, 16.0, 32.0), reg_max=16):
        super().__init__()
        self.nc = nc
        self.reg_max = reg_max
        self.no = nc + 4 * reg_max
        self.stride = torch.tensor(stride, dtype=torch.float32)

    def forward(self, x: List[torch.Tensor]) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        if self.training:
            return x
        return detect_decode(x, self.stride.tolist(), self.nc, self.reg_max), x
