"""Mirror of the decode entry points of reference core/algorithms/centernet.py: `decode_boxes`
(:271-314), `_suppress_redundant_centers` (:316-326), `_top_k` (:328-338)."""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np
import torch

from ... import ops


class CenterNetA:
    """Decode half of the reference's CenterNet algorithm class.  `cfg` is the reference's
    CenternetConfig (or any object with the same attributes)."""

    def __init__(self, cfg, device):
        self.cfg = cfg
        self.device = device
        self.num_classes = cfg.dataset.num_classes
        self.input_size = cfg.arch.input_size[1:]
        self.downsampling_ratio = cfg.arch.downsampling_ratio
        self.feature_size = [self.input_size[0] // self.downsampling_ratio,
                             self.input_size[1] // self.downsampling_ratio]
        self.K = cfg.decode.max_boxes_per_img
        self.conf_threshold = cfg.decode.score_threshold
        self.nms_threshold = cfg.decode.nms_threshold
        self.use_nms = cfg.decode.use_nms
        self.letterbox_image = cfg.decode.letterbox_image

    def decode_boxes(self, pred, h, w, conf_threshold=None):
        """pred (B, H/4, W/4, nc + 4) NHWC -> (boxes (n, 4) original-image pixels, scores (n,), classes (n,)
        int64) as numpy.  Like the reference, a batch is MERGED: rows of all images are concatenated and
        the DIoU-NMS runs over the merged set with the single (h, w) - in effect a batch-1 API."""
        if conf_threshold is None:
            conf_threshold = self.conf_threshold
        B = pred.shape[0]
        lb = ops.letterbox_params([(h, w)] * B, self.input_size, pred.device)
        if B == 1:
            det = ops.centernet_decode(pred.float(), self.K, conf_threshold, self.use_nms, self.nms_threshold, lb)
            n = int(det.count.item())
            boxes, scores, classes = det.box[0, :n], det.score[0, :n], det.cls[0, :n]
        else:
            det = ops.centernet_decode(pred.float(), self.K, conf_threshold, False, self.nms_threshold, None)
            counts = det.count.tolist()
            boxes = torch.cat([det.box[b, :n] for b, n in enumerate(counts)])
            scores = torch.cat([det.score[b, :n] for b, n in enumerate(counts)])
            classes = torch.cat([det.cls[b, :n] for b, n in enumerate(counts)])
            if self.use_nms and boxes.shape[0] > 0:
                keep = ops.diou_nms(boxes, scores, self.nms_threshold)
                boxes, scores, classes = boxes[keep], scores[keep], classes[keep]
            L = lb[0]
            boxes = boxes.clone()
            boxes[:, 0::2] = (boxes[:, 0::2] * L[0] - L[2]) * L[4]
            boxes[:, 1::2] = (boxes[:, 1::2] * L[1] - L[3]) * L[4]
        return boxes.cpu().numpy(), scores.cpu().numpy(), classes.cpu().numpy().astype(np.int64)

    def decode_batch(self, pred, image_hw: Sequence[Tuple[int, int]], conf_threshold=None) -> ops.CenterDetections:
        """Batched extension: per-image top-K, per-image DIoU-NMS, per-image letterbox; stays on device."""
        if conf_threshold is None:
            conf_threshold = self.conf_threshold
        lb = ops.letterbox_params(image_hw, self.input_size, pred.device)
        return ops.centernet_decode(pred.float(), self.K, conf_threshold, self.use_nms, self.nms_threshold, lb)

    @staticmethod
    def _top_k(scores, k):
        """(B, H, W, C) scores -> (topk_scores, inds int32, classes, ys, xs), flat index order
        (y*W + x)*C + c as in the reference (:328-338).  cvpp_topk: exact radix select, equal scores ordered
        by the lower flat index (torch.topk leaves ties unspecified) - the rule of the fused decode."""
        B, H, W, C = scores.size()
        val, _, clses, ys, xs, pixel = ops.topk(scores.reshape(B, -1).float(), k, split=(C, W))
        return val, pixel, clses, ys, xs

    @staticmethod
    def _suppress_redundant_centers(heatmap, pool_size=3):
        """heatmap * (heatmap == maxpool(heatmap)).  Applied, like the reference, to the NHWC tensor as is,
        so the window spans (x, class).  Dense helper kept for API parity (the fused
        kernel never materialises it): pool_size 3 on a CUDA tensor runs cvpp_centernet_suppress."""
        if pool_size == 3 and heatmap.is_cuda and heatmap.dim() == 4 and heatmap.dtype == torch.float32:
            return ops.centernet_suppress(heatmap)
        pad = (pool_size - 1) // 2
        hmax = torch.nn.functional.max_pool2d(heatmap, kernel_size=pool_size, stride=1, padding=pad)
        return heatmap * (heatmap == hmax).to(torch.float32)
