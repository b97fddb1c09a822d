"""Pins the C oracle against fixtures produced by the REAL reference (tests/golden/make_golden.py).

Integer outputs (kept indices, class ids, counts) must be bit-exact.  Stage-B float outputs are
bit-exact too (they are pure fp32 add/sub/mul/div on identical inputs).  Decode outputs go through
exp(), where the oracle's libm differs from torch's Sleef by <= a few ulp:
    boxes : |d| <= 1e-5 * |ref| + 4 ulp at the 640-px input scale (2.5e-4 px)
    scores: |d| <= 1e-5 * |ref|
"""
import os

import numpy as np
import pytest

import oracle
import synth

BOX_RTOL, BOX_ATOL, SCORE_RTOL = 1e-5, 2.5e-4, 1e-5


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def test_nms_known_answers(golden_dir):
    g = load(golden_dir, "nms_kat")
    for i in range(int(g["n_cases"])):
        keep = oracle.nms(g[f"boxes{i}"], g[f"scores{i}"], float(g[f"thr{i}"]))
        assert np.array_equal(keep, g[f"keep{i}"]), (i, keep, g[f"keep{i}"])


def test_nms_random_all_branches(golden_dir):
    g = load(golden_dir, "nms_random")
    for i in range(int(g["n_cases"])):
        b, s, c, t = g[f"boxes{i}"], g[f"scores{i}"], g[f"cls{i}"], float(g[f"thr{i}"])
        assert np.array_equal(oracle.nms(b, s, t), g[f"keep_nms{i}"])
        assert np.array_equal(oracle.batched_nms(b, s, c, t, 0), g[f"keep_auto{i}"])
        assert np.array_equal(oracle.batched_nms(b, s, c, t, 1), g[f"keep_trick{i}"])
        assert np.array_equal(oracle.batched_nms(b, s, c, t, 2), g[f"keep_vanilla{i}"])
        # the suppression must be non-trivial for the bigger cases
        if len(s) >= 300:
            assert len(g[f"keep_nms{i}"]) < 0.8 * len(s)


def _split(flat, counts):
    out, o = [], 0
    for c in counts:
        out.append(flat[o:o + c])
        o += c
    return out


def test_yolov8_small_decode_and_nms(golden_dir):
    g = load(golden_dir, "yolov8_small")
    levels = [g["level0"], g["level1"], g["level2"]]
    y = oracle.yolov8_decode(levels, synth.YOLOV8_STRIDES, 80)
    ref = g["y"]
    assert np.all(np.abs(y[:, :4] - ref[:, :4]) <= BOX_RTOL * np.abs(ref[:, :4]) + BOX_ATOL)
    assert np.all(np.abs(y[:, 4:] - ref[:, 4:]) <= SCORE_RTOL * np.abs(ref[:, 4:]))
    for tag in "abc":
        conf, iou, md = g[f"params_{tag}"]
        rows, anchors, _ = oracle.yolov8_nms(ref, float(conf), float(iou), int(md), nc=80)
        counts = g[f"counts_{tag}"]
        assert [len(a) for a in anchors] == list(counts)
        for r, a, rr, ra in zip(rows, anchors, _split(g[f"rows_{tag}"], counts), _split(g[f"anchors_{tag}"], counts)):
            assert np.array_equal(a, ra)
            assert np.array_equal(r, rr)          # stage B on identical fp32 inputs: bit-exact
        assert sum(counts) > 0


def test_yolov8_full_head(golden_dir):
    g = load(golden_dir, "yolov8_full")
    for tag in ("iid", "clu"):
        seed, B, clustered = [int(v) for v in g[f"{tag}_seed"]]
        levels = synth.yolov8_head(seed, B=B, clustered=bool(clustered))
        assert synth.checksum(levels) == int(g[f"{tag}_crc"]), "regenerated input drifted from the fixture"
        y = oracle.yolov8_decode(levels, synth.YOLOV8_STRIDES, 80)
        sub = g[f"{tag}_y_sub"]
        ys = y[:, :, ::53]
        assert np.all(np.abs(ys[:, :4] - sub[:, :4]) <= BOX_RTOL * np.abs(sub[:, :4]) + BOX_ATOL)
        assert np.all(np.abs(ys[:, 4:] - sub[:, 4:]) <= SCORE_RTOL * np.abs(sub[:, 4:]))
        assert np.allclose(y.astype(np.float64).sum(axis=2), g[f"{tag}_y_sum"], rtol=1e-6, atol=1e-2)
        # end to end (oracle decode -> oracle NMS) against the reference's kept sets
        for ctag, conf in (("eval", 0.001), ("pred", 0.25)):
            rows, anchors, cand = oracle.yolov8_nms(y, conf, 0.7, 300, nc=80)
            counts = g[f"{tag}_{ctag}_counts"]
            assert [len(a) for a in anchors] == list(counts)
            for r, a, rr, ra in zip(rows, anchors, _split(g[f"{tag}_{ctag}_rows"], counts),
                                    _split(g[f"{tag}_{ctag}_anchors"], counts)):
                assert np.array_equal(a, ra)
                assert np.array_equal(r[:, 5], rr[:, 5])
                assert np.all(np.abs(r[:, :4] - rr[:, :4]) <= BOX_RTOL * np.abs(rr[:, :4]) + BOX_ATOL)
                assert np.all(np.abs(r[:, 4] - rr[:, 4]) <= SCORE_RTOL * np.abs(rr[:, 4]))
            if ctag == "eval":
                assert np.all(cand > 1000)        # the vanilla batched_nms branch was exercised


def test_yolov8_pred_stage_b(golden_dir):
    g = load(golden_dir, "yolov8_pred")
    for tag in g["tags"]:
        seed, B, A, nc, nm, conf, iou, md = g[f"{tag}_cfg"]
        pred = synth.yolov8_pred(int(seed), int(B), int(A), nc=int(nc), nm=int(nm))
        assert synth.checksum([pred]) == int(g[f"{tag}_crc"])
        rows, anchors, cand = oracle.yolov8_nms(pred, float(conf), float(iou), int(md), nc=int(nc))
        counts = g[f"{tag}_counts"]
        assert [len(a) for a in anchors] == list(counts)
        for r, a, rr, ra, c in zip(rows, anchors, _split(g[f"{tag}_rows"], counts),
                                   _split(g[f"{tag}_anchors"], counts), cand):
            assert np.array_equal(a, ra)
            assert np.array_equal(r, rr[:, :6])
            assert len(a) < c                      # something was suppressed or capped


def test_centernet_decode_and_diou_nms(golden_dir):
    g = load(golden_dir, "centernet")
    for i in range(int(g["diou_cases"])):
        keep = oracle.diou_nms(g[f"diou_boxes{i}"], g[f"diou_scores{i}"], 0.5)
        assert np.array_equal(keep, g[f"diou_keep{i}"])
    assert len(g["diou_keep4"]) < 300
    for tag in g["cases"]:
        seed, B, H, W, nc, in_h, in_w = [int(v) for v in g[f"{tag}_cfg"]]
        pred = synth.centernet_pred(seed, B, H, W, nc)
        assert synth.checksum([pred]) == int(g[f"{tag}_crc"])
        lb = oracle.letterbox_params([(480, 640)] * B, (in_h, in_w))
        for ntag, use_nms in (("nms", True), ("raw", False)):
            for ctag, conf in (("lo", 0.001), ("hi", 0.1)):
                out = oracle.centernet_decode(pred, 100, conf, 0, use_nms, 0.5, lb)
                for b, (box, score, cls, pix) in enumerate(out):
                    k = f"{tag}_{b}_{ntag}_{ctag}"
                    assert np.array_equal(cls, g[k + "_classes"])
                    assert np.all(np.abs(score - g[k + "_scores"]) <= SCORE_RTOL * g[k + "_scores"])
                    assert np.all(np.abs(box - g[k + "_boxes"]) <= BOX_RTOL * np.abs(g[k + "_boxes"]) + BOX_ATOL)
                    assert len(cls) > 0


def test_ssd_decode(golden_dir):
    g = load(golden_dir, "ssd")
    loc, conf = synth.ssd_head(int(g["seed"][0]), int(g["seed"][1]))
    assert synth.checksum([loc, conf]) == int(g["crc"])
    pri = oracle.ssd_priors()
    assert np.array_equal(pri, load(golden_dir, "host_helpers")["ssd_priors"])
    for ctag, thr in (("eval", 0.001), ("pred", 0.7)):
        out = oracle.ssd_decode(loc, conf, pri, thr, 0.5)
        for b, (rows, prior) in enumerate(out):
            ref = g[f"rows_{b}_{ctag}"]
            got = oracle.yolo_correct_rows(rows, (300, 300), (480, 640), True)
            assert got.shape == ref.shape and np.array_equal(got[:, 4], ref[:, 4])
            assert np.all(np.abs(got[:, 5] - ref[:, 5]) <= SCORE_RTOL * ref[:, 5])
            assert np.all(np.abs(got[:, :4] - ref[:, :4]) <= BOX_RTOL * np.abs(ref[:, :4]) + BOX_ATOL * 2.2)


def _close(got, ref, rtol, atol=0.0):
    return got.shape == ref.shape and bool(np.all(np.abs(got - ref) <= rtol * np.abs(ref) + atol))


def test_yolov7_decode_and_nms(golden_dir):
    g = load(golden_dir, "yolov7")
    for tag in g["cases"]:
        seed, B, nc = [int(v) for v in g[f"{tag}_cfg"]]
        levels = synth.yolov7_head(seed, B, nc=nc)
        assert synth.checksum(levels) == int(g[f"{tag}_crc"])
        dec = oracle.yolov7_decode(levels, nc)
        for ctag, thr in (("eval", 0.001), ("pred", 0.5)):
            out, cand = oracle.yolov7_nms(dec, thr, 0.3)
            for b, (rows, anchors) in enumerate(out):
                ref = g[f"{tag}_{b}_{ctag}"]
                got = oracle.yolo_correct_rows(rows, (640, 640), (480, 640), True)
                assert got.shape == ref.shape and np.array_equal(got[:, 6], ref[:, 6])      # class ids, order
                assert _close(got[:, 4:6], ref[:, 4:6], SCORE_RTOL)
                assert _close(got[:, :4], ref[:, :4], BOX_RTOL, BOX_ATOL)
                assert (len(rows) == 0) == bool(g[f"{tag}_{b}_{ctag}_none"])
                if ctag == "eval":
                    assert 0 < len(rows) < cand[b]                                          # NMS suppressed something
        if tag == "voc":
            assert synth.checksum([dec]) == int(g["voc_decoded_crc"])
            out, _ = oracle.yolov7_nms(dec, 0.01, 0.4)
            for b, (rows, anchors) in enumerate(out):
                ref = g[f"voc_{b}_free"]
                got = oracle.yolo_correct_rows(rows, (640, 640), (333, 500), False)
                assert got.shape == ref.shape and np.array_equal(got[:, 4:], ref[:, 4:])   # same decoded input: exact
                assert _close(got[:, :4], ref[:, :4], 1e-6, 1e-5)


def test_yolov3_decoder_and_nms(golden_dir):
    g = load(golden_dir, "yolov3")
    for tag in ("one", "two"):
        seed, B, merged = [int(v) for v in g[f"{tag}_cfg"]]
        levels = synth.yolov3_head(seed, B, merged=bool(merged))
        assert synth.checksum(levels) == int(g[f"{tag}_crc"])
        boxes, scores = oracle.yolov3_dense(levels, 20)
        for ctag, thr in (("eval", 0.001), ("pred", 0.6)):
            ob, os_, oc, orow, cands = oracle.yolo3_nms(boxes, scores, thr, 0.5)
            assert np.array_equal(oc, g[f"{tag}_{ctag}_classes"]) and oc.dtype == np.int32
            assert _close(os_, g[f"{tag}_{ctag}_scores"], SCORE_RTOL)
            assert _close(ob, g[f"{tag}_{ctag}_boxes"], BOX_RTOL, 1e-6)
            assert 0 < len(oc) < cands
    # dense predict_bounding_bbox on the 13 x 13 scale
    levels = synth.yolov3_head(51, 1)
    b13, s13 = oracle.yolov3_dense(levels[:1], 20)
    xy, wh = g["pbb_xy"].reshape(-1, 2), g["pbb_wh"].reshape(-1, 2)
    ref_boxes = np.concatenate([xy - wh / 2, xy + wh / 2], 1)
    assert _close(b13, ref_boxes, BOX_RTOL, 1e-6)
    ref_scores = (g["pbb_conf"].reshape(-1, 1) * np.ones((1, 20), np.float32)).reshape(-1)[::7]
    # scores = conf * prob: compare through the sub-sampled prob digest
    assert _close(s13.reshape(-1)[::7], ref_scores * g["pbb_prob_sub"], SCORE_RTOL)
    # standalone yolo3_nms on identical inputs: exact
    ob, os_, oc, orow, cands = oracle.yolo3_nms(g["nms3_boxes_in"], g["nms3_scores_in"], 0.5, 0.45)
    assert np.array_equal(ob, g["nms3_boxes"]) and np.array_equal(os_[:, None], g["nms3_scores"])
    assert np.array_equal(oc, g["nms3_classes"]) and len(oc) < cands
