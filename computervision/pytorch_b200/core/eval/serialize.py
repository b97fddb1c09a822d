"""Detection serialisation in the reference's two evaluation formats.

* VOC txt lines  (core/algorithms/yolo_v8.py:286-296 and its twins in yolo_v7.py / ssd.py / centernet.py):
  ``f"{class_name} {str(score)[:6]} {int(left)} {int(top)} {int(right)} {int(bottom)}\\n"`` per detection,
  one file per image, detections in the order the decoder returned them.
* COCO json rows (yolo_v8.py:364-372): ``{"image_id", "category_id", "bbox": [x, y, w, h], "score"}``.

The box truncation / xywh conversion happen on the device (cvpp_detection_epilogue layouts CVPP_ROWS_VOC /
CVPP_ROWS_COCO); what is left for the host is string formatting of rows that arrive in ONE transfer per batch.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np


def _per_image(rows: np.ndarray, counts: Sequence[int]):
    """rows either padded (B, max_det, 6) or compact (sum(counts), 6) -> the (n_b, 6) block of every image."""
    if rows.ndim == 3:
        for b, n in enumerate(counts):
            yield rows[b, :n]
    else:
        o = 0
        for n in counts:
            yield rows[o:o + n]
            o += n


def voc_lines(rows: np.ndarray, counts: Sequence[int], class_names: Sequence[str], pad_empty: bool = False) -> List[List[str]]:
    """rows float32 in CVPP_ROWS_VOC layout [cls, score, l, t, r, b] (box already truncated on the device), padded
    (B, max_det, 6) or compact (sum(counts), 6) -> per image the list of lines the reference writes.
    `str(np.float32)` gives the shortest repr exactly like `str(scores[i])` in the reference.  An image without
    detections: YOLOv8 writes nothing (yolo_v8.py:278-283 pads the boxes but iterates over the empty class list);
    YOLOv7 / SSD / CenterNet pad the whole result with one all-zero row, i.e. write "<class 0> 0.0 0 0 0 0"
    (yolo_v7.py:128-130, ssd.py:130-132, centernet.py:171-175) - `pad_empty=True`."""
    out = []
    for block in _per_image(rows, counts):
        lines = [f"{class_names[int(r[0])]} {str(np.float32(r[1]))[:6]} {int(r[2])} {int(r[3])} {int(r[4])} {int(r[5])}\n"
                 for r in block]
        if pad_empty and not lines:
            lines = [f"{class_names[0]} {str(np.float32(0.0))[:6]} 0 0 0 0\n"]
        out.append(lines)
    return out


def coco_results(rows: np.ndarray, counts: Sequence[int], image_ids: Sequence[int],
                 clsid2catid: Sequence[int]) -> List[Dict]:
    """rows float32 in CVPP_ROWS_COCO layout [x, y, w, h, score, cls], padded (B, max_det, 6) or compact
    (sum(counts), 6) -> the reference's `results` list (one dict per detection, images in order)."""
    res = []
    for b, block in enumerate(_per_image(rows, counts)):
        for r in block:
            res.append({"image_id": int(image_ids[b]), "category_id": clsid2catid[int(r[5])],
                        "bbox": [float(r[0]), float(r[1]), float(r[2]), float(r[3])], "score": float(r[4])})
    return res
