"""Mirror of the decode entry points of reference core/algorithms/ssd.py: `decode_boxes` (:236-288),
`_parse_mbox_loc` (:290-325), `_get_ssd_anchors` (:482-541)."""
from __future__ import annotations

import numpy as np
import torch

from ... import ops
from ..utils.anchor import generate_ssd_anchor_v2


class Ssd:
    """Decode half of the reference's SSD algorithm class.  `cfg` is the reference's SsdConfig (or any
    object with the same attributes)."""

    def __init__(self, cfg, device):
        self.cfg = cfg
        self.device = device
        self.input_image_size = cfg.arch.input_size[1:]
        self.num_classes = cfg.dataset.num_classes
        self.anchor_sizes = cfg.arch.anchor_sizes
        self.feature_shapes = cfg.arch.feature_shapes
        self.aspect_ratios = cfg.arch.aspect_ratios
        self.anchors = self._get_ssd_anchors()
        self.num_anchors = self.anchors.shape[0]
        self.variance = np.repeat(np.array(cfg.loss.variance, dtype=np.float32), 2, axis=0)
        self.letterbox_image = cfg.decode.letterbox_image
        self.conf_threshold = cfg.decode.conf_threshold
        self.nms_threshold = cfg.decode.nms_threshold
        self._priors_dev = None
        if not np.allclose(self.variance[::2], [0.1, 0.2]):
            raise NotImplementedError("libcvpp compiles the reference's variance (0.1, 0.2) into the SSD decoder")

    def _get_ssd_anchors(self):
        return generate_ssd_anchor_v2(self.input_image_size, self.anchor_sizes, self.feature_shapes, self.aspect_ratios)

    def _priors(self, device):
        if self._priors_dev is None or self._priors_dev.device != device:
            self._priors_dev = torch.from_numpy(self.anchors).to(device)   # uploaded once, not per call (:292)
        return self._priors_dev

    def _parse_mbox_loc(self, mbox_loc):
        """(8732, 4) regression output of one image -> decoded, clamped, normalised xyxy."""
        return ops.ssd_parse_loc(mbox_loc.float(), self._priors(mbox_loc.device))

    def decode_boxes(self, preds, h, w, conf_threshold=None):
        """preds = (loc (B, 8732, 4), conf logits (B, 8732, nc + 1)) -> list of B float32 ndarrays (n_i, 6)
        [x1, y1, x2, y2, label, conf] in original-image pixels, class-ascending then score-descending;
        an image without detections yields [] like the reference."""
        if conf_threshold is None:
            conf_threshold = self.conf_threshold
        loc, conf = preds[0].float(), preds[1].float()
        cand = ops.ssd_decode_filter(loc, conf, self._priors(loc.device), conf_threshold,
                                     max_cand=min(self.num_anchors * self.num_classes, 32768))
        try:
            det = ops.per_class_nms_device(cand, self.nms_threshold)
        except OverflowError:
            cand = ops.ssd_decode_filter(loc, conf, self._priors(loc.device), conf_threshold)
            det = ops.per_class_nms_device(cand, self.nms_threshold)
        B = int(loc.shape[0])
        table = ops.correct_boxes_params([(h, w)] * B, self.input_image_size, self.letterbox_image, loc.device)
        rows = ops.detection_epilogue(det, ops.ROWS_SSD, ops.BOX_CORRECT, table)   # ssd.py:275-287 on the device
        rows_h, counts = rows.cpu().numpy(), det.count.cpu().tolist()               # one transfer for the batch
        return [rows_h[b, :n].copy() if n > 0 else [] for b, n in enumerate(counts)]
