"""-m gpu: serialisation epilogue + batched evaluation driver against the
reference's own per-image path: YOLOv8.decode_box (mirror) followed by the formatting code of
evaluate_on_voc / evaluate_on_coco (core/algorithms/yolo_v8.py:286-296, :364-372) restated inline."""
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

import synth

pytestmark = pytest.mark.gpu

from computervision.pytorch_b200.core.algorithms.yolo_v8 import YOLOv8  # noqa: E402
from computervision.pytorch_b200.core.eval import BatchedDetectionEvaluator  # noqa: E402
from computervision.pytorch_b200.core.models.yolov8.modules import detect_decode  # noqa: E402

DEV = "cuda:0"
NAMES = [f"c{i}" for i in range(80)]


def _cfg():
    return NS(arch=NS(input_size=(3, 640, 640)), dataset=NS(num_classes=80),
              decode=NS(conf_threshold=0.25, nms_threshold=0.7, max_det=300, letterbox_image=True))


def test_voc_lines_and_coco_rows_match_the_per_image_reference_path():
    levels = [torch.from_numpy(l).to(DEV) for l in synth.yolov8_head(77, B=5, clustered=True)]
    hw = [(480, 640), (1080, 1920), (640, 427), (333, 500), (640, 640)]
    ev = BatchedDetectionEvaluator(80, synth.YOLOV8_STRIDES, (640, 640), batch_size=5)
    lines = ev.voc_batch(levels, hw, NAMES)
    catid = list(range(1, 81))
    coco = ev.coco_batch(levels, hw, [100 + i for i in range(5)], catid)
    algo = YOLOv8(_cfg(), DEV)
    y = detect_decode(levels, synth.YOLOV8_STRIDES, 80)
    want_lines, want_coco = [], []
    for b, (h, w) in enumerate(hw):
        boxes, scores, cls = algo.decode_box(y[b:b + 1], h, w, conf_threshold=0.001)   # reference per-image path
        want_lines.append([f"{NAMES[int(c)]} {str(scores[i])[:6]} {int(boxes[i, 0])} {int(boxes[i, 1])} {int(boxes[i, 2])} {int(boxes[i, 3])}\n"
                           for i, c in enumerate(cls)])
        for i, c in enumerate(cls):
            l, t, r, bt = boxes[i]
            want_coco.append({"image_id": 100 + b, "category_id": catid[c], "bbox": [float(l), float(t), float(r - l), float(bt - t)],
                              "score": float(scores[i])})
    assert lines == want_lines
    assert coco == want_coco
    assert sum(len(x) for x in lines) > 1000


def test_evaluate_coco_driver_batches_and_preserves_order():
    heads = synth.yolov8_head(78, B=7)
    dev_levels = [torch.from_numpy(l).to(DEV) for l in heads]
    hw = [(480, 640)] * 7

    def head_fn(idx):
        return [l[idx] for l in dev_levels], [hw[i] for i in idx]

    catid = list(range(1, 81))
    a = BatchedDetectionEvaluator(80, synth.YOLOV8_STRIDES, (640, 640), batch_size=3).evaluate_coco(7, head_fn, list(range(7)), catid)
    b = BatchedDetectionEvaluator(80, synth.YOLOV8_STRIDES, (640, 640), batch_size=7).evaluate_coco(7, head_fn, list(range(7)), catid)
    assert a == b and [r["image_id"] for r in a] == sorted(r["image_id"] for r in a)
