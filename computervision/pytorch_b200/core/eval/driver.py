"""Batched, image-sharded replacement of the bs=1 loops in the reference's `evaluate_on_voc` /
`evaluate_on_coco` (core/algorithms/yolo_v8.py:244-372 and the twins yolo_v7.py:94-231, ssd.py:96-233,
centernet.py:137-268): the reference decodes one image per iteration (`assert len(preds) == 1`, yolo_v8.py:229)
with a device->host sync per image (per class for SSD, ssd.py:278); here a whole batch goes through
decode+filter -> fused sort+NMS -> compact epilogue on the device and comes back in ONE pinned-memory transfer
([rows | row offsets | flags] in a single buffer), and with several ranks every rank handles a contiguous shard
of the image list (no collective on the data path; the per-rank result lists are merged once at the end).

One evaluator class per head family, all sharing the sharding, the staging and the serialisation:

    BatchedDetectionEvaluator   YOLOv8   head_fn -> (levels [(B,144,H,W)...], image_hw)
    YOLOv7Evaluator             YOLOv7   head_fn -> (levels [(B,255,H,W)...], image_hw)
    SsdEvaluator                SSD      head_fn -> ((loc (B,P,4), conf (B,P,nc+1)), image_hw)
    CenterNetEvaluator          CenterNet head_fn -> (pred (B,H,W,nc+4), image_hw)
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from ... import ops
from ...distributed import shard_range
from .serialize import coco_results, voc_lines


class _RowsEvaluator:
    """Sharding + one-shot pinned staging + serialisation.  Subclasses implement `_detect`."""

    pad_empty_voc = True   # the reference pads an empty result with one all-zero detection of class 0 (ssd.py:130-132,
                           # yolo_v7.py:128-130, centernet.py:171-175); YOLOv8 alone writes an empty file

    def __init__(self, input_hw: Sequence[int], letterbox_image: bool = True, batch_size: int = 64):
        self.input_hw = tuple(int(v) for v in input_hw)
        self.letterbox_image = bool(letterbox_image)
        self.batch_size = int(batch_size)
        self._dev_buf: Optional[torch.Tensor] = None
        self._host_buf: Optional[torch.Tensor] = None

    # -- sharding ---------------------------------------------------------------------------------
    @staticmethod
    def my_indices(n_images: int) -> range:
        if dist.is_available() and dist.is_initialized():
            s, e = shard_range(n_images, dist.get_world_size(), dist.get_rank())
            return range(s, e)
        return range(n_images)

    def batches(self, n_images: int) -> Iterable[List[int]]:
        idx = list(self.my_indices(n_images))
        for i in range(0, len(idx), self.batch_size):
            yield idx[i:i + self.batch_size]

    # -- one batch ----------------------------------------------------------------------------------
    def _detect(self, head, image_hw):
        """-> (detections on the device, box_mode, letterbox table, aux_dense, candidates or None)."""
        raise NotImplementedError

    def _grow(self, what: str) -> bool:
        """Enlarge the capacity that overflowed ("cand" / "out"); False when there is nothing left to grow."""
        return False

    def _rows(self, head, image_hw: Sequence[Tuple[int, int]], layout: int):
        """-> (rows (n_total, 6) float32 ndarray, counts list): ONE device->host transfer per batch, from one device
        buffer [rows | row_offset (B+1) | overflow | raw detection counts (B) | raw candidate counts (B)] into pinned
        host memory.  A capacity that turns out too small (the raw counts say so) is grown and the batch redone."""
        while True:
            det, box_mode, table, aux, cand = self._detect(head, image_hw)
            B, max_out = int(det.box.shape[0]), int(det.box.shape[1])
            dev = det.box.device
            cap = B * max_out
            n = cap * 6 + 3 * B + 2
            if self._dev_buf is None or self._dev_buf.numel() < n or self._dev_buf.device != dev:
                self._dev_buf = torch.empty((n,), dtype=torch.float32, device=dev)
                self._host_buf = torch.empty((n,), dtype=torch.float32).pin_memory()
            buf = self._dev_buf[:n]
            tail = cap * 6
            off = buf[tail: tail + B + 1].view(torch.int32)
            ovf = buf[tail + B + 1: tail + B + 2].view(torch.int32)
            buf[tail + B + 2: tail + 2 * B + 2].view(torch.int32).copy_(det.count)
            if cand is not None:
                buf[tail + 2 * B + 2:].view(torch.int32).copy_(cand.count)
            else:
                buf[tail + 2 * B + 2:].zero_()
            ops.detection_epilogue_compact(det, layout, cap, box_mode, table, aux, out=buf[:tail].view(cap, 6),
                                           row_offset=off, overflow=ovf)
            host = self._host_buf[:n]
            host.copy_(buf, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()        # the one transfer of the batch has landed
            h = host.numpy()
            offsets = h[tail: tail + B + 1].view(np.int32)
            raw_det = h[tail + B + 2: tail + 2 * B + 2].view(np.int32)
            raw_cand = h[tail + 2 * B + 2:].view(np.int32)
            if cand is not None and int(raw_cand.max(initial=0)) > cand.max_cand:
                if not self._grow("cand"):
                    raise OverflowError(f"{int(raw_cand.max())} candidates in one image exceed the buffer ({cand.max_cand})")
                continue
            if int(raw_det.max(initial=0)) > max_out:
                if not self._grow("out"):
                    raise OverflowError(f"{int(raw_det.max())} detections in one image exceed the output capacity {max_out}")
                continue
            total = int(offsets[B])
            return h[: total * 6].reshape(total, 6).copy(), np.diff(offsets).tolist()

    def voc_batch(self, head, image_hw, class_names) -> List[List[str]]:
        rows, counts = self._rows(head, image_hw, ops.ROWS_VOC)
        return voc_lines(rows, counts, class_names, pad_empty=self.pad_empty_voc)

    def coco_batch(self, head, image_hw, image_ids, clsid2catid) -> List[Dict]:
        rows, counts = self._rows(head, image_hw, ops.ROWS_COCO)
        return coco_results(rows, counts, image_ids, clsid2catid)

    # -- whole data set -----------------------------------------------------------------------------
    @staticmethod
    def _merge(mine: list) -> list:
        """Shards are contiguous image ranges in rank order, so concatenating the per-rank lists in rank order
        restores dataset order; one all_gather_object at the end (host objects, not on the data path)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            parts: List[Optional[list]] = [None] * dist.get_world_size()
            dist.all_gather_object(parts, mine)
            return [r for part in parts for r in part]
        return mine

    def evaluate_coco(self, n_images: int, head_fn: Callable, image_ids: Sequence[int], clsid2catid: Sequence[int]) -> List[Dict]:
        """The reference's `results` list (evaluate_on_coco: yolo_v8.py:345-372 and twins), identical on every rank."""
        mine: List[Dict] = []
        for idx in self.batches(n_images):
            head, image_hw = head_fn(idx)
            mine += self.coco_batch(head, image_hw, [image_ids[i] for i in idx], clsid2catid)
        return self._merge(mine)

    def evaluate_voc(self, n_images: int, head_fn: Callable, class_names: Sequence[str]) -> List[List[str]]:
        """Per image (dataset order) the lines of its `detection-results/<image_id>.txt` (evaluate_on_voc:
        yolo_v8.py:266-296 and twins), identical on every rank."""
        mine: List[List[str]] = []
        for idx in self.batches(n_images):
            head, image_hw = head_fn(idx)
            mine += self.voc_batch(head, image_hw, class_names)
        return self._merge(mine)


class BatchedDetectionEvaluator(_RowsEvaluator):
    """YOLOv8-head evaluator.

    head_fn(batch_index_list) -> (levels, image_hw): the per-level head tensors (B, 4*16 + nc, H, W) on the
    device for those dataset indices (i.e. `model(images)` of the reference loop before the Detect tail) and
    the original (h, w) of every image.  The evaluator owns the sharding, the post-processing and the
    serialisation."""

    pad_empty_voc = False   # yolo_v8.py:278-283: only `results[0]` is padded, the class list stays empty

    def __init__(self, nc: int, strides: Sequence[float], input_hw: Sequence[int], letterbox_image: bool = True,
                 conf_threshold: float = 0.001, iou_threshold: float = 0.7, max_det: int = 300, batch_size: int = 64):
        super().__init__(input_hw, letterbox_image, batch_size)
        self.nc, self.strides = int(nc), tuple(float(s) for s in strides)
        self.conf, self.iou, self.max_det = float(conf_threshold), float(iou_threshold), int(max_det)
        self._post: Optional[ops.Yolov8Postprocessor] = None

    def _detect(self, levels, image_hw):
        ls = ops.make_levels(list(levels), self.strides)
        if self._post is None or self._post.B != ls.B or self._post.A != ls.A or self._post.device != ls.device:
            self._post = ops.Yolov8Postprocessor(ls.B, ls.A, self.nc, ls.device, max_det=self.max_det)
        det = self._post(ls, self.conf, self.iou)
        table = ops.correct_boxes_params(image_hw, self.input_hw, self.letterbox_image, ls.device)
        return det, ops.BOX_NORMALISE_CORRECT, table, None, None


class YOLOv7Evaluator(_RowsEvaluator):
    """YOLOv7 (reference yolo_v7.py:94-231): 3-level anchor decode, obj * class_conf >= conf, per-class NMS without a
    cap, rows score = obj * class_conf (yolo_v7.py:133,212).  `anchors` (9, 2) pixels, `anchors_mask` as in the cfg."""

    def __init__(self, nc: int, anchors, anchors_mask, input_hw: Sequence[int], letterbox_image: bool = True,
                 conf_threshold: float = 0.001, nms_threshold: float = 0.3, batch_size: int = 64, max_out: int = 4096):
        super().__init__(input_hw, letterbox_image, batch_size)
        self.nc = int(nc)
        a = np.asarray(anchors, dtype=np.float32).reshape(-1, 2)
        self.level_anchors = np.concatenate([a[list(m)] for m in anchors_mask], axis=0)
        self.conf, self.nms, self.max_out = float(conf_threshold), float(nms_threshold), int(max_out)

    def _detect(self, levels, image_hw):
        ls = ops.make_levels([l.float() for l in levels])
        cand = ops.yolov7_decode_filter(ls, self.nc, self.level_anchors[: 3 * ls.n], self.input_hw, self.conf)
        det = ops.sort_nms(cand, self.nms, ops.RULE_PER_CLASS, ops.ORDER_CLASS_MAJOR, max_det=0,
                           max_out=min(self.max_out, cand.max_cand))
        table = ops.correct_boxes_params(image_hw, self.input_hw, self.letterbox_image, ls.device)
        return det, ops.BOX_CORRECT, table, None, cand

    def _grow(self, what: str) -> bool:
        if what == "out" and self.max_out < 3 * 25200:
            self.max_out *= 4
            return True
        return False


class SsdEvaluator(_RowsEvaluator):
    """SSD (reference ssd.py:96-233): prior decode + softmax, per-(prior, class) filter, per-class NMS."""

    def __init__(self, nc: int, priors: np.ndarray, input_hw: Sequence[int], letterbox_image: bool = True,
                 conf_threshold: float = 0.001, nms_threshold: float = 0.5, batch_size: int = 64, max_out: int = 4096):
        super().__init__(input_hw, letterbox_image, batch_size)
        self.nc, self.priors_np = int(nc), np.ascontiguousarray(priors, dtype=np.float32)
        self.conf, self.nms, self.max_out = float(conf_threshold), float(nms_threshold), int(max_out)
        self.max_cand = 32768
        self._priors: Optional[torch.Tensor] = None

    def _detect(self, head, image_hw):
        loc, conf = head[0].float(), head[1].float()
        if self._priors is None or self._priors.device != loc.device:
            self._priors = torch.from_numpy(self.priors_np).to(loc.device)
        P = int(loc.shape[1])
        self._full_cand = P * self.nc
        cand = ops.ssd_decode_filter(loc, conf, self._priors, self.conf, max_cand=min(self._full_cand, self.max_cand))
        det = ops.sort_nms(cand, self.nms, ops.RULE_PER_CLASS, ops.ORDER_CLASS_MAJOR, max_det=0,
                           max_out=min(self.max_out, cand.max_cand))
        table = ops.correct_boxes_params(image_hw, self.input_hw, self.letterbox_image, loc.device)
        return det, ops.BOX_CORRECT, table, None, cand

    def _grow(self, what: str) -> bool:
        if what == "cand" and self.max_cand < self._full_cand:
            self.max_cand = self._full_cand           # dense conf maps: every (prior, class) pair may pass
            return True
        if what == "out" and self.max_out < self._full_cand:
            self.max_out *= 4
            return True
        return False


class CenterNetEvaluator(_RowsEvaluator):
    """CenterNet (reference centernet.py:137-268): peak extraction + top-K (+ DIoU-NMS) with the per-image letterbox
    inverse fused into the decode kernel."""

    def __init__(self, input_hw: Sequence[int], K: int = 100, conf_threshold: float = 0.001, use_nms: bool = False,
                 nms_threshold: float = 0.5, batch_size: int = 64):
        super().__init__(input_hw, True, batch_size)
        self.K, self.conf, self.use_nms, self.nms = int(K), float(conf_threshold), bool(use_nms), float(nms_threshold)

    def _detect(self, pred, image_hw):
        lb = ops.letterbox_params(image_hw, self.input_hw, pred.device)
        d = ops.centernet_decode(pred.float(), self.K, self.conf, self.use_nms, self.nms, lb)
        det = ops.Detections(box=d.box, score=d.score, cls=d.cls, anchor=d.pixel, count=d.count)
        return det, ops.BOX_KEEP, None, None, None
