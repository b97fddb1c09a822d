"""-m gpu parity tests of the YOLOv8 path: CUDA kernels (through the C ABI) vs the CPU oracle and the
committed reference fixtures.

Contract (SURVEY.md §7 hard part 4, §8c):
  stage A  head -> candidates: same candidate anchors and classes as the oracle; boxes/scores within
           tolerance (GPU expf vs CPU libm differ by ulps);
  stage B  candidates -> kept set on IDENTICAL fp32 inputs: kept anchors, classes, order, boxes and
           scores bit-exact;
  end to end: kept anchors/classes equal to the reference's on the committed seeds.
"""
import os

import numpy as np
import pytest
import torch

import oracle
import synth
from gpu_util import (assert_boxes_close, assert_scores_close, candidates_to_numpy, dets_to_numpy, to_dev)

pytestmark = pytest.mark.gpu

from computervision.pytorch_b200 import ops  # noqa: E402

DEV = "cuda:0"
NC = 80


def _split(flat, counts):
    out, o = [], 0
    for c in counts:
        out.append(flat[o:o + c])
        o += c
    return out


def _run_e2e(levels_np, conf, iou, max_det=300, rule=ops.RULE_TORCHVISION_CPU):
    lv = to_dev(levels_np)
    ls = ops.make_levels(lv, synth.YOLOV8_STRIDES)
    post = ops.Yolov8Postprocessor(ls.B, ls.A, NC, DEV, max_det=max_det)
    det = post(ls, conf, iou, rule=rule)
    torch.cuda.synchronize()
    return dets_to_numpy(det), det.cand_count.cpu().numpy()


@pytest.mark.parametrize("clustered", [False, True])
@pytest.mark.parametrize("conf", [0.001, 0.25])
def test_stage_a_candidates_match_oracle(clustered, conf):
    levels = synth.yolov8_head(1000 + int(clustered), B=3, clustered=clustered)
    y = oracle.yolov8_decode(levels, synth.YOLOV8_STRIDES, NC)
    ref = oracle.yolov8_candidates(y, conf, nc=NC)
    ls = ops.make_levels(to_dev(levels), synth.YOLOV8_STRIDES)
    got = candidates_to_numpy(ops.yolov8_decode_filter(ls, NC, conf))
    for g, (rbox, rscore, rcls, ranchor) in zip(got, ref):
        assert np.array_equal(g["anchor"], ranchor)
        assert np.array_equal(g["cls"], rcls)
        assert_scores_close(g["score"], rscore)
        assert_boxes_close(g["box"], rbox)
        assert len(ranchor) > 0


def test_decode_full_matches_oracle_and_filter_kernel():
    levels = synth.yolov8_head(77, B=2, clustered=True)
    y_ref = oracle.yolov8_decode(levels, synth.YOLOV8_STRIDES, NC)
    ls = ops.make_levels(to_dev(levels), synth.YOLOV8_STRIDES)
    y = ops.yolov8_decode_full(ls, NC).cpu().numpy()
    assert_boxes_close(y[:, :4], y_ref[:, :4])
    assert_scores_close(y[:, 4:], y_ref[:, 4:])
    # the fused filter kernel and the dense kernel must agree bit for bit on what they both compute
    cands = candidates_to_numpy(ops.yolov8_decode_filter(ls, NC, 0.001))
    ref = oracle.yolov8_candidates(y, 0.001, nc=NC)   # oracle filter applied to the GPU's own y
    for g, (rbox, rscore, rcls, ranchor) in zip(cands, ref):
        assert np.array_equal(g["anchor"], ranchor)
        assert np.array_equal(g["cls"], rcls)
        assert np.array_equal(g["score"], rscore)
        assert np.array_equal(g["box"], rbox)


@pytest.mark.parametrize("tag", ["p0", "p1", "p2", "p3"])
def test_stage_b_bit_exact_vs_reference_fixture(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "yolov8_pred.npz"))
    seed, B, A, nc, nm, conf, iou, md = g[f"{tag}_cfg"]
    B, A, nc, nm, md = int(B), int(A), int(nc), int(nm), int(md)
    pred = synth.yolov8_pred(int(seed), B, A, nc=nc, nm=nm)
    assert synth.checksum([pred]) == int(g[f"{tag}_crc"])
    c = ops.pred_filter(torch.from_numpy(pred).to(DEV), nc, float(conf))
    ops.segmented_sort(c, ops.RULE_TORCHVISION_CPU, max_nms=30000)
    det = dets_to_numpy(ops.nms(c, float(iou), ops.RULE_TORCHVISION_CPU, ops.ORDER_SCORE_DESC, max_det=md))
    counts = g[f"{tag}_counts"]
    assert [len(d["anchor"]) for d in det] == list(counts)
    for d, rr, ra in zip(det, _split(g[f"{tag}_rows"], counts), _split(g[f"{tag}_anchors"], counts)):
        assert np.array_equal(d["anchor"], ra)
        assert np.array_equal(d["cls"], rr[:, 5].astype(np.int32))
        assert np.array_equal(d["score"], rr[:, 4])
        assert np.array_equal(d["box"], rr[:, :4])


@pytest.mark.parametrize("rule", [ops.RULE_TORCHVISION_CPU, ops.RULE_COORD_TRICK, ops.RULE_PER_CLASS])
@pytest.mark.parametrize("conf,max_det", [(0.001, 300), (0.25, 300), (0.05, 100000)])
def test_stage_b_bit_exact_vs_oracle_all_rules(rule, conf, max_det):
    pred = synth.yolov8_pred(99, 4, 4200, nc=NC, n_clusters=60, per_cluster=20, background=900)
    rows, anchors, cand = oracle.yolov8_nms(pred, conf, 0.6, max_det, nc=NC, nms_mode=rule)
    c = ops.pred_filter(torch.from_numpy(pred).to(DEV), NC, conf)
    ops.segmented_sort(c, rule, max_nms=30000)
    det = dets_to_numpy(ops.nms(c, 0.6, rule, ops.ORDER_SCORE_DESC, max_det=max_det, max_out=min(max_det, 4200)))
    assert np.array_equal(c.count.cpu().numpy(), cand)
    for d, r, a in zip(det, rows, anchors):
        assert np.array_equal(d["anchor"], a)
        assert np.array_equal(d["cls"], r[:, 5].astype(np.int32))
        assert np.array_equal(d["score"], r[:, 4])
        assert np.array_equal(d["box"], r[:, :4])
    assert sum(len(a) for a in anchors) < int(cand.sum())


@pytest.mark.parametrize("tag", ["iid", "clu"])
@pytest.mark.parametrize("ctag,conf", [("eval", 0.001), ("pred", 0.25)])
def test_end_to_end_vs_reference_fixture(golden_dir, tag, ctag, conf):
    g = np.load(os.path.join(golden_dir, "yolov8_full.npz"))
    seed, B, clustered = [int(v) for v in g[f"{tag}_seed"]]
    levels = synth.yolov8_head(seed, B=B, clustered=bool(clustered))
    assert synth.checksum(levels) == int(g[f"{tag}_crc"])
    det, _ = _run_e2e(levels, conf, 0.7)
    counts = g[f"{tag}_{ctag}_counts"]
    assert [len(d["anchor"]) for d in det] == list(counts)
    for d, rr, ra in zip(det, _split(g[f"{tag}_{ctag}_rows"], counts), _split(g[f"{tag}_{ctag}_anchors"], counts)):
        assert np.array_equal(d["anchor"], ra)
        assert np.array_equal(d["cls"], rr[:, 5].astype(np.int32))
        assert_scores_close(d["score"], rr[:, 4])
        assert_boxes_close(d["box"], rr[:, :4])


def test_end_to_end_c2_batch64_vs_oracle():
    """BASELINE config 2 at full size: bs=64, conf .001, IoU .7, max_det 300."""
    levels = synth.yolov8_head(2024, B=64, clustered=False)
    det, cand = _run_e2e(levels, 0.001, 0.7)
    y = oracle.yolov8_decode(levels, synth.YOLOV8_STRIDES, NC)
    rows, anchors, cand_ref = oracle.yolov8_nms(y, 0.001, 0.7, 300, nc=NC)
    assert np.array_equal(cand, cand_ref)
    assert cand.min() > 1000
    mism = 0
    for d, r, a in zip(det, rows, anchors):
        if not (np.array_equal(d["anchor"], a) and np.array_equal(d["cls"], r[:, 5].astype(np.int32))):
            mism += 1
            continue
        assert_scores_close(d["score"], r[:, 4])
        assert_boxes_close(d["box"], r[:, :4])
    assert mism == 0, f"{mism}/64 images differ from the oracle end to end"


def test_end_to_end_is_stage_b_exact_on_own_decode():
    """NMS on the GPU's own decoded tensor must equal the oracle's NMS on that same tensor exactly."""
    levels = synth.yolov8_head(555, B=8, clustered=True)
    ls = ops.make_levels(to_dev(levels), synth.YOLOV8_STRIDES)
    y = ops.yolov8_decode_full(ls, NC).cpu().numpy()
    for conf in (0.001, 0.25):
        det, cand = _run_e2e(levels, conf, 0.7)
        rows, anchors, cand_ref = oracle.yolov8_nms(y, conf, 0.7, 300, nc=NC)
        assert np.array_equal(cand, cand_ref)
        for d, r, a in zip(det, rows, anchors):
            assert np.array_equal(d["anchor"], a)
            assert np.array_equal(d["cls"], r[:, 5].astype(np.int32))
            assert np.array_equal(d["score"], r[:, 4])
            assert np.array_equal(d["box"], r[:, :4])


def test_generic_kernel_equals_tma_kernel_and_handles_ragged_levels(monkeypatch):
    levels = synth.yolov8_head(31, B=2, clustered=True)
    ls = ops.make_levels(to_dev(levels), synth.YOLOV8_STRIDES)
    a = candidates_to_numpy(ops.yolov8_decode_filter(ls, NC, 0.01))
    monkeypatch.setenv("CVPP_FORCE_GENERIC", "1")
    b = candidates_to_numpy(ops.yolov8_decode_filter(ls, NC, 0.01))
    monkeypatch.delenv("CVPP_FORCE_GENERIC")
    for x, y in zip(a, b):
        for k in ("anchor", "cls", "score", "box"):
            assert np.array_equal(x[k], y[k])
    # ragged: level sizes that are not multiples of 4 cells (bulk copies impossible -> generic kernel)
    sizes = ((15, 13), (7, 9), (3, 5))
    lv = synth.yolov8_head(32, B=2, sizes=sizes, cls_mu=-5.0, cls_sigma=3.0)
    y = oracle.yolov8_decode(lv, synth.YOLOV8_STRIDES, NC)
    ref = oracle.yolov8_candidates(y, 0.01, nc=NC)
    got = candidates_to_numpy(ops.yolov8_decode_filter(ops.make_levels(to_dev(lv), synth.YOLOV8_STRIDES), NC, 0.01))
    for g, (rbox, rscore, rcls, ranchor) in zip(got, ref):
        assert np.array_equal(g["anchor"], ranchor) and np.array_equal(g["cls"], rcls)
        assert_scores_close(g["score"], rscore)
        assert_boxes_close(g["box"], rbox)


def test_x_cat_views_and_partial_tiles():
    """Levels given as slices of the concatenated (B, C, A) tensor (Detect's x_cat, modules.py:438)."""
    sizes = ((20, 12), (10, 6), (5, 4))        # 240 / 60 / 20 cells: partial 128-cell tiles everywhere
    lv = synth.yolov8_head(8, B=3, sizes=sizes, cls_mu=-5.0, cls_sigma=3.0)
    x_cat = np.concatenate([l.reshape(3, 144, -1) for l in lv], axis=2)
    t = torch.from_numpy(x_cat).to(DEV)
    views, off = [], 0
    for h, w in sizes:
        views.append(t[:, :, off:off + h * w])
        off += h * w
    ls = ops.make_levels(views, synth.YOLOV8_STRIDES, sizes=sizes)
    got = candidates_to_numpy(ops.yolov8_decode_filter(ls, NC, 0.01))
    ref = oracle.yolov8_candidates(oracle.yolov8_decode(lv, synth.YOLOV8_STRIDES, NC), 0.01, nc=NC)
    for g, (rbox, rscore, rcls, ranchor) in zip(got, ref):
        assert np.array_equal(g["anchor"], ranchor) and np.array_equal(g["cls"], rcls)
        assert_boxes_close(g["box"], rbox)


def test_edge_cases_empty_and_ties():
    # nothing passes the threshold -> zero detections everywhere
    levels = synth.yolov8_head(3, B=2, clustered=False)
    det, cand = _run_e2e(levels, 1.0, 0.7)
    assert cand.tolist() == [0, 0] and all(len(d["anchor"]) == 0 for d in det)
    # equal scores: lower anchor index first, duplicates keep the lower index (SURVEY §4 known answers)
    A, nc = 64, 4
    pred = np.zeros((1, 4 + nc, A), np.float32)
    pred[0, 0:2] = 100.0
    pred[0, 2:4] = 10.0
    for a, (cx, s, c) in {5: (100.0, 0.7, 1), 9: (100.0, 0.7, 1), 20: (300.0, 0.7, 1), 33: (300.0, 0.7, 2)}.items():
        pred[0, 0, a] = cx
        pred[0, 4 + c, a] = s
    rows, anchors, _ = oracle.yolov8_nms(pred, 0.25, 0.5, 300, nc=nc)
    c = ops.pred_filter(torch.from_numpy(pred).to(DEV), nc, 0.25)
    ops.segmented_sort(c)
    d = dets_to_numpy(ops.nms(c, 0.5))[0]
    assert d["anchor"].tolist() == anchors[0].tolist() == [5, 20, 33]
    assert np.array_equal(d["box"], rows[0][:, :4])


def test_properties_full_size():
    """Size-independent properties at BASELINE config-2 size (bs=64)."""
    levels = synth.yolov8_head(4242, B=64, clustered=True)
    ls = ops.make_levels(to_dev(levels), synth.YOLOV8_STRIDES)
    cands = candidates_to_numpy(ops.yolov8_decode_filter(ls, NC, 0.001))
    post = ops.Yolov8Postprocessor(ls.B, ls.A, NC, DEV, max_det=300)
    det = dets_to_numpy(post(ls, 0.001, 0.7))
    for c, d in zip(cands, det):
        assert len(d["anchor"]) <= 300
        assert np.all(np.diff(d["score"]) <= 0)                         # score-descending
        assert len(np.unique(d["anchor"])) == len(d["anchor"])
        pos = np.searchsorted(c["anchor"], d["anchor"])
        assert np.array_equal(c["anchor"][pos], d["anchor"])             # kept is a subset of candidates
        assert np.array_equal(c["cls"][pos], d["cls"]) and np.array_equal(c["score"][pos], d["score"])
        # idempotence: NMS over the kept boxes alone keeps all of them
        keep = oracle.batched_nms(d["box"], d["score"], d["cls"].astype(np.float32), 0.7, mode=2)
        assert len(keep) == len(d["anchor"])


def test_pipelined_postprocess_equals_serial():
    """Throughput mode (two buffer sets / graphs / streams, batches in flight): every slot of every submit holds
    exactly the detections of the serial call."""
    levels = synth.yolov8_head(31, B=16, clustered=True)
    ls = ops.make_levels(to_dev(levels), synth.YOLOV8_STRIDES)
    A = ls.A
    ref = ops.Yolov8Postprocessor(16, A, NC, torch.device(DEV))(ls, 0.001, 0.7)
    torch.cuda.synchronize()
    want = [t.clone() for t in (ref.box, ref.score, ref.cls, ref.anchor, ref.count)]
    for graph in (True, False):
        pipe = ops.PipelinedPostprocess(16, A, NC, torch.device(DEV), ls, 0.001, 0.7, depth=2, graph=graph)
        pipe.fork()
        dets = [pipe.submit() for _ in range(7)]
        pipe.join()
        torch.cuda.synchronize()
        assert dets[0] is dets[2] and dets[0] is not dets[1]          # slots alternate
        for d in dets[-2:]:
            n = d.count.cpu().numpy()
            assert np.array_equal(n, want[4].cpu().numpy()) and n.sum() > 0
            for got, w in zip((d.box, d.score, d.cls, d.anchor), want[:4]):
                for b in range(16):
                    assert torch.equal(got[b, :n[b]], w[b, :n[b]])


def test_pipelined_postprocess_distinct_batches_with_a_producer():
    """One input buffer set per slot, refilled by a producer stream with a DIFFERENT batch every step: the producer
    waits on `consumed[slot]` before overwriting a slot's inputs and hands `ready` to submit().  Every step's
    detections must equal the serial result of ITS batch (a missing handshake shows as detections of a
    neighbouring batch)."""
    dev = torch.device(DEV)
    B, steps, depth = 8, 9, 3
    batches = [to_dev(synth.yolov8_head(100 + k, B=B, clustered=bool(k & 1))) for k in range(4)]
    serial = ops.Yolov8Postprocessor(B, 8400, NC, dev)
    want = []
    for lv in batches:
        d = serial(ops.make_levels(lv, synth.YOLOV8_STRIDES), 0.001, 0.7)
        torch.cuda.synchronize()
        want.append((d.count.clone(), d.anchor.clone(), d.cls.clone(), d.box.clone()))
    assert not torch.equal(want[0][1], want[1][1])
    for graph in (True, False):
        slots_in = [[torch.empty_like(t) for t in batches[0]] for _ in range(depth)]
        pipe = ops.PipelinedPostprocess(B, 8400, NC, dev, [ops.make_levels(s, synth.YOLOV8_STRIDES) for s in slots_in],
                                        0.001, 0.7, graph=graph)
        assert not pipe.static_input and len(pipe.streams) == depth
        producer = torch.cuda.Stream(device=dev)
        got = []
        for k in range(steps):
            slot = pipe.next_slot
            with torch.cuda.stream(producer):
                producer.wait_event(pipe.consumed[slot])
                for dst, src in zip(slots_in[slot], batches[k % len(batches)]):
                    dst.copy_(src, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(producer)
            det = pipe.submit(ready)
            # snapshot the slot's detections on its own stream before the slot is reused
            with torch.cuda.stream(pipe.streams[slot]):
                got.append((det.count.clone(), det.anchor.clone(), det.cls.clone(), det.box.clone()))
        pipe.join()
        torch.cuda.synchronize()
        for k, g in enumerate(got):
            w = want[k % len(batches)]
            assert torch.equal(g[0], w[0]), (graph, k)
            for b, n in enumerate(w[0].tolist()):
                assert torch.equal(g[1][b, :n], w[1][b, :n]) and torch.equal(g[2][b, :n], w[2][b, :n])
                assert torch.equal(g[3][b, :n], w[3][b, :n])


def test_postprocessor_rejects_small_candidate_buffers_and_foreign_devices():
    with pytest.raises(ValueError):
        ops.Yolov8Postprocessor(2, 8400, NC, torch.device(DEV), max_cand=1000)
    levels = synth.yolov8_head(3, B=2)
    ls = ops.make_levels(to_dev(levels), synth.YOLOV8_STRIDES)
    post = ops.Yolov8Postprocessor(2, 8400, NC, torch.device(DEV))
    post.max_cand = 100                                   # bypass the Python check: the C ABI must refuse too
    from computervision.pytorch_b200._lib import CvppError
    with pytest.raises(CvppError):
        post(ls, 0.001, 0.7)
