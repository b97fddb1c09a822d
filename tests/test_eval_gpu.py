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


# ------------------------------------------------------------------------------------------------------------------
# every head family against the ORACLE (the C restatement of the reference) + the reference's formatting restated
# ------------------------------------------------------------------------------------------------------------------
import os  # noqa: E402
import subprocess  # noqa: E402
import sys  # noqa: E402

import oracle  # noqa: E402
from computervision.pytorch_b200.core.eval import CenterNetEvaluator, SsdEvaluator, YOLOv7Evaluator  # noqa: E402

HW = [(480, 640), (1080, 1920), (640, 427), (333, 500)]


def _ref_voc_lines(boxes, scores, cls, names, pad):
    """evaluate_on_voc's formatting (yolo_v8.py:286-296 and twins) on one image's oracle result."""
    if len(cls) == 0 and pad:
        boxes, scores, cls = np.zeros((1, 4), np.float32), np.zeros((1,), np.float32), np.zeros((1,), np.int32)
    return [f"{names[int(c)]} {str(scores[i])[:6]} {int(boxes[i, 0])} {int(boxes[i, 1])} {int(boxes[i, 2])} {int(boxes[i, 3])}\n"
            for i, c in enumerate(cls)]


def _assert_lines_close(got, want):
    """Same detections in the same order; the oracle's libm differs from the GPU's exp by a few ulp, so a coordinate
    within 1e-3 px of an integer may truncate differently and the 6-character score may differ in its last digit."""
    assert len(got) == len(want)
    n = off = 0
    for gl, wl in zip(got, want):
        assert len(gl) == len(wl), (len(gl), len(wl))
        for g, w in zip(gl, wl):
            gt, wt = g.split(), w.split()
            assert gt[0] == wt[0]
            assert abs(float(gt[1]) - float(wt[1])) <= 2e-4 * max(abs(float(wt[1])), 1e-3)
            d = [abs(int(a) - int(b)) for a, b in zip(gt[2:], wt[2:])]
            assert max(d) <= 1
            off += sum(d)
            n += 4
    assert n > 0 and off <= max(2, n // 500), (off, n)


def test_yolov8_evaluator_vs_oracle():
    heads = synth.yolov8_head(91, B=4, clustered=True)
    ev = BatchedDetectionEvaluator(80, synth.YOLOV8_STRIDES, (640, 640), batch_size=4)
    lines = ev.voc_batch([torch.from_numpy(l).to(DEV) for l in heads], HW, NAMES)
    rows, _, _ = oracle.yolov8_nms(oracle.yolov8_decode(heads, synth.YOLOV8_STRIDES, 80), 0.001, 0.7, 300, nc=80)
    want = []
    for b, r in enumerate(rows):
        r = r.copy()
        r[:, 0:4:2] /= 640
        r[:, 1:4:2] /= 640                                                      # yolo_v8.py:233-234
        r = oracle.yolo_correct_rows(r, (640, 640), HW[b], True)
        want.append(_ref_voc_lines(r[:, :4], r[:, 4], r[:, 5].astype(np.int32), NAMES, pad=False))
    _assert_lines_close(lines, want)


def test_yolov7_evaluator_vs_oracle_and_empty_image_padding():
    levels = synth.yolov7_head(92, B=4)
    levels[0][3, :] = -20.0
    levels[1][3, :] = -20.0
    levels[2][3, :] = -20.0                                                       # image 3: no candidate at all
    ev = YOLOv7Evaluator(80, oracle.YOLOV7_ANCHORS, oracle.YOLOV7_MASK, (640, 640), conf_threshold=0.001, nms_threshold=0.3,
                         batch_size=4, max_out=512)                             # 512 < kept/img: exercises the regrow
    dev_levels = [torch.from_numpy(l).to(DEV) for l in levels]
    lines = ev.voc_batch(dev_levels, HW, NAMES)
    assert ev.max_out > 512
    res, _ = oracle.yolov7_nms(oracle.yolov7_decode(levels, 80), 0.001, 0.3)
    want = []
    for b, (r, _a) in enumerate(res):
        r = oracle.yolo_correct_rows(r, (640, 640), HW[b], True)
        want.append(_ref_voc_lines(r[:, :4], r[:, 4] * r[:, 5], r[:, 6].astype(np.int32), NAMES, pad=True))
    assert want[3] == ["c0 0.0 0 0 0 0\n"]
    _assert_lines_close(lines, want)
    coco = ev.coco_batch(dev_levels, HW, [7, 8, 9, 10], list(range(1, 81)))
    assert len(coco) == sum(len(r) for r, _ in res) and {c["image_id"] for c in coco} == {7, 8, 9}
    r0 = oracle.yolo_correct_rows(res[0][0], (640, 640), HW[0], True)
    assert coco[0]["category_id"] == int(r0[0, 6]) + 1
    assert np.allclose(coco[0]["bbox"], [r0[0, 0], r0[0, 1], r0[0, 2] - r0[0, 0], r0[0, 3] - r0[0, 1]], rtol=1e-4, atol=2e-3)


def test_ssd_evaluator_vs_oracle():
    loc, conf = synth.ssd_head(93, 4)
    names = [f"v{i}" for i in range(20)]
    ev = SsdEvaluator(20, oracle.ssd_priors(), (300, 300), conf_threshold=0.001, nms_threshold=0.5, batch_size=4)
    lines = ev.voc_batch((torch.from_numpy(loc).to(DEV), torch.from_numpy(conf).to(DEV)), HW, names)
    res = oracle.ssd_decode(loc, conf, oracle.ssd_priors(), 0.001, 0.5)
    want = []
    for b, (r, _p) in enumerate(res):
        r = oracle.yolo_correct_rows(r, (300, 300), HW[b], True)
        want.append(_ref_voc_lines(r[:, :4], r[:, 5], r[:, 4].astype(np.int32), names, pad=True))
    _assert_lines_close(lines, want)


def test_centernet_evaluator_vs_oracle():
    pred = synth.centernet_pred(94, 4, 96, 96, 20)
    names = [f"v{i}" for i in range(20)]
    for use_nms in (False, True):
        ev = CenterNetEvaluator((384, 384), K=100, conf_threshold=0.001, use_nms=use_nms, batch_size=4)
        lines = ev.voc_batch(torch.from_numpy(pred).to(DEV), HW, names)
        res = oracle.centernet_decode(pred, 100, 0.001, 0, use_nms, 0.5, oracle.letterbox_params(HW, (384, 384)))
        want = [_ref_voc_lines(bx, sc, cl, names, pad=True) for bx, sc, cl, _ in res]
        _assert_lines_close(lines, want)


def test_evaluators_sharded_over_two_ranks_reproduce_the_single_process_lists():
    """world size 2 (gloo for the host-side merge, both ranks on this GPU): contiguous image shards, per-rank
    evaluation, one all_gather_object - the merged VOC / COCO lists equal the single-process ones on every rank."""
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_eval_world2.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", script],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "world2 ok" in r.stdout
