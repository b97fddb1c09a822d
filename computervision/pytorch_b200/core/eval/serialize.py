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


def voc_lines(rows: np.ndarray, counts: Sequence[int], class_names: Sequence[str]) -> List[List[str]]:
    """rows (B, max_det, 6) float32 in CVPP_ROWS_VOC layout [cls, score, l, t, r, b] (box already truncated on
    the device) -> per image the list of lines the reference writes.  `str(np.float32)` gives the shortest
    repr exactly like `str(scores[i])` in the reference; an image without detections gets no line (the
    reference then substitutes a zero box but iterates over an empty class list, so it writes nothing)."""
    out = []
    for b, n in enumerate(counts):
        lines = []
        for r in rows[b, :n]:
            lines.append(f"{class_names[int(r[0])]} {str(np.float32(r[1]))[:6]} {int(r[2])} {int(r[3])} {int(r[4])} {int(r[5])}\n")
        out.append(lines)
    return out


def coco_results(rows: np.ndarray, counts: Sequence[int], image_ids: Sequence[int],
                 clsid2catid: Sequence[int]) -> List[Dict]:
    """rows (B, max_det, 6) float32 in CVPP_ROWS_COCO layout [x, y, w, h, score, cls] -> the reference's
    `results` list (one dict per detection, images in order)."""
    res = []
    for b, n in enumerate(counts):
        for r in rows[b, :n]:
            res.append({"image_id": int(image_ids[b]), "category_id": clsid2catid[int(r[5])],
                        "bbox": [float(r[0]), float(r[1]), float(r[2]), float(r[3])], "score": float(r[4])})
    return res
