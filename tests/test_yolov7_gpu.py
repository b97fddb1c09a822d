"""-m gpu: YOLOv7 anchor decode + per-class NMS vs the oracle (bit-exact anchors / classes) and the
reference fixtures; includes BASELINE config 5's per-GPU shard shape (25 200 anchors x 85)."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

import oracle
import synth
from gpu_util import record_error, BOX_ATOL, BOX_RTOL, SCORE_RTOL, decode_keys

pytestmark = pytest.mark.gpu

from computervision.pytorch_b200 import ops  # noqa: E402
from computervision.pytorch_b200.core.algorithms.yolo_v7 import YOLOv7  # noqa: E402
from computervision.pytorch_b200.core.utils.nms import yolo7_nms  # noqa: E402

DEV = "cuda:0"
ANCHORS = [12, 16, 19, 36, 40, 28, 36, 75, 76, 55, 72, 146, 142, 110, 192, 243, 459, 401]


def _cfg(nc):
    return NS(arch=NS(input_size=(3, 640, 640), anchors=ANCHORS, anchors_mask=[[6, 7, 8], [3, 4, 5], [0, 1, 2]]),
              dataset=NS(num_classes=nc), decode=NS(letterbox_image=True, conf_threshold=0.5, nms_threshold=0.3))


def _close(got, ref, rtol, atol=0.0):
    record_error(got, ref, "box" if atol else "score_or_norm_box")
    return got.shape == ref.shape and bool(np.all(np.abs(got - ref) <= rtol * np.abs(ref) + atol))


def test_decode_box_vs_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "yolov7.npz"))
    for tag in g["cases"]:
        seed, B, nc = [int(v) for v in g[f"{tag}_cfg"]]
        levels = synth.yolov7_head(seed, B, nc=nc)
        assert synth.checksum(levels) == int(g[f"{tag}_crc"])
        algo = YOLOv7(_cfg(nc), DEV)
        preds = [torch.from_numpy(l).to(DEV) for l in levels]
        for ctag, thr in (("eval", 0.001), ("pred", None)):
            res = algo.decode_box(preds, 480, 640, thr)
            for b, r in enumerate(res):
                ref = g[f"{tag}_{b}_{ctag}"]
                if bool(g[f"{tag}_{b}_{ctag}_none"]):
                    assert r is None
                    continue
                assert r.dtype == np.float32 and r.shape == ref.shape
                assert np.array_equal(r[:, 6], ref[:, 6])                         # class ids in class-major order
                assert _close(r[:, 4:6], ref[:, 4:6], SCORE_RTOL)
                assert _close(r[:, :4], ref[:, :4], BOX_RTOL, BOX_ATOL)
        res = algo.decode_box(preds, 480, 640, 0.9999999)
        assert [r is None for r in res] == [bool(v) for v in g[f"{tag}_empty_is_none"]]
        if tag == "voc":
            dec = torch.from_numpy(oracle.yolov7_decode(levels, nc)).to(DEV)
            res = yolo7_nms(dec, nc, [640, 640], [333, 500], False, DEV, conf_thres=0.01, nms_thres=0.4)
            res2 = algo._nms(dec, [640, 640], [333, 500], 0.01)                   # same path, nms_threshold 0.3
            for b, r in enumerate(res):
                ref = g[f"voc_{b}_free"]
                assert r.shape == ref.shape and np.array_equal(r[:, 4:], ref[:, 4:])   # identical decoded input
                assert _close(r[:, :4], ref[:, :4], 1e-6, 1e-5)
                assert res2[b] is not None and res2[b].shape[1] == 7


@pytest.mark.parametrize("B,nc,conf", [(8, 80, 0.001), (3, 20, 0.5), (2, 80, 0.0)])
def test_candidates_and_kept_anchors_exact_vs_oracle(B, nc, conf):
    """C5 shard shape: every candidate (anchor, class) and every kept anchor index equals the oracle's."""
    levels = synth.yolov7_head(100 + B, B, nc=nc)
    dec = oracle.yolov7_decode(levels, nc)
    out, cand_ref = oracle.yolov7_nms(dec, conf, 0.3)
    ls = ops.make_levels([torch.from_numpy(l).to(DEV) for l in levels])
    cand = ops.yolov7_decode_filter(ls, nc, oracle.yolov7_level_anchors(), (640, 640), conf)
    cnt = cand.count.cpu().numpy()
    assert np.array_equal(cnt, cand_ref)
    key = cand.key.cpu().numpy().view(np.uint64)
    aux = cand.aux_dense.cpu().numpy()
    dense = cand.box_dense.cpu().numpy()
    if B <= 3:                                                                    # candidate-level check
        for b in range(B):
            cls, score, anchor = decode_keys(key[b, :cnt[b]])
            o = np.argsort(anchor)
            d = dec[b, anchor[o]]
            assert np.array_equal(cls[o], d[:, 5:].argmax(1))
            assert _close(aux[b, anchor[o], 0], d[:, 4], SCORE_RTOL) and _close(aux[b, anchor[o], 1], d[:, 5:].max(1), SCORE_RTOL)
            xyxy = np.concatenate([d[:, :2] - d[:, 2:4] / 2, d[:, :2] + d[:, 2:4] / 2], 1)
            assert _close(dense[b, anchor[o]], xyxy, BOX_RTOL, 1e-6)
    det = ops.per_class_nms_device(cand, 0.3)
    n = det.count.cpu().numpy()
    for b, (rows, anchors) in enumerate(out):
        assert n[b] == len(anchors)
        assert np.array_equal(det.anchor[b, :n[b]].cpu().numpy(), anchors), b     # kept anchors: bit-exact
        assert np.array_equal(det.cls[b, :n[b]].cpu().numpy(), rows[:, 6].astype(np.int32))
        assert _close(det.box[b, :n[b]].cpu().numpy(), rows[:, :4], BOX_RTOL, 1e-6)
    if 0 < conf < 0.01:
        assert int(n.sum()) < int(cnt.sum())


def test_generic_kernel_matches_stream_kernel(monkeypatch):
    """Unaligned level views take the thread-per-anchor kernel; results must equal the TMA kernel's."""
    levels = synth.yolov7_head(5, 2, nc=20)
    ls = ops.make_levels([torch.from_numpy(l).to(DEV) for l in levels])
    a = ops.yolov7_decode_filter(ls, 20, oracle.yolov7_level_anchors(), (640, 640), 0.001)
    monkeypatch.setenv("CVPP_FORCE_GENERIC", "1")
    b = ops.yolov7_decode_filter(ls, 20, oracle.yolov7_level_anchors(), (640, 640), 0.001)
    assert torch.equal(a.count, b.count)
    for i in range(2):
        n = int(a.count[i])
        ka, kb = a.key[i, :n].sort().values, b.key[i, :n].sort().values
        assert torch.equal(ka, kb)
        anc = (ka & 0x1FFFFF).long()
        assert torch.equal(a.box_dense[i, anc], b.box_dense[i, anc]) and torch.equal(a.aux_dense[i, anc], b.aux_dense[i, anc])


def test_pred_filter_stage_exact_on_identical_inputs():
    """yolov7_pred_filter + sort + NMS on a decoded tensor equals the oracle's _nms bit for bit."""
    levels = synth.yolov7_head(9, 3, nc=80)
    dec = oracle.yolov7_decode(levels, 80)
    out, cand_ref = oracle.yolov7_nms(dec, 0.001, 0.3)
    cand = ops.yolov7_pred_filter(torch.from_numpy(dec).to(DEV), 80, 0.001)
    assert np.array_equal(cand.count.cpu().numpy(), cand_ref)
    det = ops.per_class_nms_device(cand, 0.3)
    rows = ops.detection_epilogue(det, ops.ROWS_YOLOV7, ops.BOX_KEEP, None, cand.aux_dense).cpu().numpy()
    n = det.count.cpu().numpy()
    for b, (ref, anchors) in enumerate(out):
        assert np.array_equal(det.anchor[b, :n[b]].cpu().numpy(), anchors)
        assert np.array_equal(rows[b, :n[b]], ref)                                # pure fp32 add/sub/mul: exact
        assert not rows[b, n[b]:].any()


def test_epilogue_matches_reference_box_correction():
    """cvpp_detection_epilogue's letterbox inverse == the numpy arithmetic of yolo_correct_boxes, bit for bit."""
    levels = synth.yolov7_head(11, 2, nc=20)
    dec = oracle.yolov7_decode(levels, 20)
    out, _ = oracle.yolov7_nms(dec, 0.01, 0.3)
    cand = ops.yolov7_pred_filter(torch.from_numpy(dec).to(DEV), 20, 0.01)
    det = ops.per_class_nms_device(cand, 0.3)
    n = det.count.cpu().numpy()
    for letterbox, hw in ((True, [(480, 640), (1080, 1920)]), (False, [(333, 500), (640, 427)])):
        table = ops.correct_boxes_params(hw, (640, 640), letterbox, DEV)
        rows = ops.detection_epilogue(det, ops.ROWS_YOLOV7, ops.BOX_CORRECT, table, cand.aux_dense).cpu().numpy()
        for b, (ref, _) in enumerate(out):
            want = oracle.yolo_correct_rows(ref, (640, 640), hw[b], letterbox)
            assert np.array_equal(rows[b, :n[b]], want)


def test_c5_shard_full_size_greedy_nms_invariants():
    """BASELINE config 5 per-GPU shard (128 images x 25 200 anchors x 85, eval threshold) through size-independent
    properties.  With distinct scores the greedy per-class NMS result is the UNIQUE set such that (a) no two kept
    boxes of a class overlap above the threshold and (b) every dropped candidate is suppressed by a kept box of its
    class with a higher score - both are checked with numpy in torchvision's fp32 arithmetic on sampled images;
    ordering / membership properties are checked on all 128."""
    B, nc, thr = 128, 80, 0.3
    g = torch.Generator(device=DEV).manual_seed(5)
    levels = []
    for s in (20, 40, 80):
        x = torch.randn((B, 3, 5 + nc, s, s), generator=g, device=DEV)
        x[:, :, 4] = x[:, :, 4] * 3.0 - 9.0
        x[:, :, 5:] = x[:, :, 5:] * 2.0 - 1.0
        levels.append(x.reshape(B, 3 * (5 + nc), s, s))
    ls = ops.make_levels(levels)
    cand = ops.yolov7_decode_filter(ls, nc, oracle.yolov7_level_anchors(), (640, 640), 0.001)
    det = ops.per_class_nms_device(cand, thr)
    cnt, n = cand.count.cpu().numpy(), det.count.cpu().numpy()
    assert cnt.min() > 3000 and np.all(n <= cnt) and np.all(n > 0)
    key = cand.key.cpu().numpy().view(np.uint64)
    dense = cand.box_dense.cpu().numpy()
    d_anchor, d_cls, d_score = det.anchor.cpu().numpy(), det.cls.cpu().numpy(), det.score.cpu().numpy()

    def iou_gt(b1, b2):  # torchvision nms_kernel.cpp arithmetic: float32 ops, compare as double
        xx1, yy1 = np.maximum(b1[:, None, 0], b2[None, :, 0]), np.maximum(b1[:, None, 1], b2[None, :, 1])
        xx2, yy2 = np.minimum(b1[:, None, 2], b2[None, :, 2]), np.minimum(b1[:, None, 3], b2[None, :, 3])
        w, h = np.maximum(np.float32(0), xx2 - xx1), np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
        a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (a1[:, None] + a2[None, :] - inter)
        return ovr.astype(np.float64) > thr

    for b in range(B):
        k = n[b]
        # class-major order, scores descending inside a class, every kept anchor is a candidate of that class
        order_key = d_cls[b, :k].astype(np.int64) * (1 << 32) - d_score[b, :k].view(np.int32).astype(np.int64)
        assert np.all(np.diff(order_key) >= 0), b      # (random data: exact score ties are possible)
    for b in (0, 17, 63, 127):
        c_cls, c_score, c_anchor = decode_keys(key[b, :cnt[b]])
        kept = set(zip(d_cls[b, :n[b]].tolist(), d_anchor[b, :n[b]].tolist()))
        assert kept <= set(zip(c_cls.tolist(), c_anchor.tolist()))
        suppressed_total = 0
        for c in np.unique(c_cls):
            m = c_cls == c
            a, s = c_anchor[m], c_score[m]
            is_kept = np.array([(int(c), int(x)) in kept for x in a])
            kb, ks = dense[b, a[is_kept]], s[is_kept]
            if is_kept.sum() > 1:                                   # (a) kept boxes are mutually compatible
                ov = iou_gt(kb, kb)
                np.fill_diagonal(ov, False)
                assert not ov.any(), (b, int(c))
            if (~is_kept).any():                                    # (b) every dropped box has a better kept suppressor
                db, ds = dense[b, a[~is_kept]], s[~is_kept]
                ov = iou_gt(db, kb) & (ks[None, :] >= ds[:, None])
                assert ov.any(axis=1).all(), (b, int(c))
                suppressed_total += int((~is_kept).sum())
        assert suppressed_total == cnt[b] - n[b] and suppressed_total > 0
