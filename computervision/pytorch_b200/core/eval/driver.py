"""Batched, image-sharded replacement of the bs=1 loops in the reference's `evaluate_on_voc` /
`evaluate_on_coco` (core/algorithms/yolo_v8.py:266-296, :345-372 and twins): the reference decodes one image
per iteration (`assert len(preds) == 1`, yolo_v8.py:229) with a device->host sync per image; here a whole batch
goes through decode+filter -> fused sort+NMS -> epilogue on the device and comes back in one transfer, and
with several ranks every rank handles a contiguous shard of the image list (no collective on the data path;
the per-rank result lists are merged once at the end)."""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from ... import ops
from ...distributed import shard_range
from .serialize import coco_results, voc_lines


class BatchedDetectionEvaluator:
    """YOLOv8-head evaluator.

    head_fn(batch_index_list) -> (levels, image_hw): the per-level head tensors (B, 4*16 + nc, H, W) on the
    device for those dataset indices (i.e. `model(images)` of the reference loop before the Detect tail) and
    the original (h, w) of every image.  The evaluator owns the sharding, the post-processing and the
    serialisation."""

    def __init__(self, nc: int, strides: Sequence[float], input_hw: Sequence[int], letterbox_image: bool = True,
                 conf_threshold: float = 0.001, iou_threshold: float = 0.7, max_det: int = 300, batch_size: int = 64):
        self.nc, self.strides, self.input_hw = int(nc), tuple(float(s) for s in strides), tuple(input_hw)
        self.letterbox_image = bool(letterbox_image)
        self.conf, self.iou, self.max_det, self.batch_size = float(conf_threshold), float(iou_threshold), int(max_det), int(batch_size)
        self._post: Optional[ops.Yolov8Postprocessor] = None

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
    def _rows(self, levels: Sequence[torch.Tensor], image_hw: Sequence[Tuple[int, int]], layout: int):
        ls = ops.make_levels(list(levels), self.strides)
        if self._post is None or self._post.B != ls.B or self._post.A != ls.A:
            self._post = ops.Yolov8Postprocessor(ls.B, ls.A, self.nc, ls.device, max_det=self.max_det)
        det = self._post(ls, self.conf, self.iou)
        table = ops.correct_boxes_params(image_hw, self.input_hw, self.letterbox_image, ls.device)
        packed = ops.detection_epilogue(det, layout, ops.BOX_NORMALISE_CORRECT, table, packed=True)
        host = packed.cpu().numpy()                              # the one device->host transfer of the batch
        n_rows = ls.B * self.max_det * 6
        return host[:n_rows].reshape(ls.B, self.max_det, 6), host[n_rows:].astype(int).tolist()

    def voc_batch(self, levels, image_hw, class_names) -> List[List[str]]:
        rows, counts = self._rows(levels, image_hw, ops.ROWS_VOC)
        return voc_lines(rows, counts, class_names)

    def coco_batch(self, levels, image_hw, image_ids, clsid2catid) -> List[Dict]:
        rows, counts = self._rows(levels, image_hw, ops.ROWS_COCO)
        return coco_results(rows, counts, image_ids, clsid2catid)

    # -- whole data set -----------------------------------------------------------------------------
    def evaluate_coco(self, n_images: int, head_fn: Callable, image_ids: Sequence[int], clsid2catid: Sequence[int]) -> List[Dict]:
        """Returns the full `results` list (identical on every rank): every rank serialises its shard, the
        shards are concatenated in rank order (= dataset order) with one all_gather_object at the end."""
        mine: List[Dict] = []
        for idx in self.batches(n_images):
            levels, image_hw = head_fn(idx)
            mine += self.coco_batch(levels, image_hw, [image_ids[i] for i in idx], clsid2catid)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            parts: List[Optional[List[Dict]]] = [None] * dist.get_world_size()
            dist.all_gather_object(parts, mine)
            return [r for part in parts for r in part]
        return mine
