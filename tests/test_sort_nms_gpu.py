"""-m gpu: the fused sort+NMS kernel (cvpp_sort_nms) must reproduce the two-kernel path
(cvpp_segmented_sort + cvpp_nms) bit for bit, in every regime: coordinate-trick images (<= 1000
candidates), per-class images, classes larger than one warp sort, images that do not fit shared memory
(fallback inside the call), max_nms truncation, both output orders, empty images."""
import numpy as np
import pytest
import torch

import oracle
import synth

pytestmark = pytest.mark.gpu

from computervision.pytorch_b200 import ops  # noqa: E402

DEV = "cuda:0"


def _both(cand_fn, iou, rule, order, max_det, max_nms=0, max_out=None):
    a = cand_fn()
    ops.segmented_sort(a, rule, max_nms=max_nms)
    ref = ops.nms(a, iou, rule, order, max_det=max_det, max_out=max_out)
    got = ops.sort_nms(cand_fn(), iou, rule, order, max_det=max_det, max_nms=max_nms, max_out=max_out)
    torch.cuda.synchronize()
    assert torch.equal(got.count, ref.count)
    cap = ref.box.shape[1]
    for b, n in enumerate(ref.count.tolist()):
        n = min(n, cap)
        assert torch.equal(got.anchor[b, :n], ref.anchor[b, :n]), b
        assert torch.equal(got.cls[b, :n], ref.cls[b, :n])
        assert torch.equal(got.score[b, :n], ref.score[b, :n]) and torch.equal(got.box[b, :n], ref.box[b, :n])
    return ref


@pytest.mark.parametrize("conf,iou,max_det", [(0.001, 0.7, 300), (0.25, 0.7, 300), (0.001, 0.45, 20), (0.05, 0.6, 1000)])
def test_yolov8_score_order_all_rules(conf, iou, max_det):
    pred = torch.from_numpy(synth.yolov8_pred(21, 6, 8400, nc=80)).to(DEV)
    pred[5, 4:] = 0.0                                                         # an image without candidates
    for rule in (ops.RULE_TORCHVISION_CPU, ops.RULE_COORD_TRICK, ops.RULE_PER_CLASS):
        ref = _both(lambda: ops.pred_filter(pred, 80, conf), iou, rule, ops.ORDER_SCORE_DESC, max_det, max_nms=30000)
        assert int(ref.count[5]) == 0 and int(ref.count[:5].min()) > 0


def test_vs_oracle_kept_anchors_exact():
    pred = synth.yolov8_pred(33, 4, 8400, nc=80)
    rows, anchors, _ = oracle.yolov8_nms(pred, 0.001, 0.7, 300, nc=80)
    det = ops.sort_nms(ops.pred_filter(torch.from_numpy(pred).to(DEV), 80, 0.001), 0.7, max_det=300, max_nms=30000)
    for b in range(4):
        n = int(det.count[b])
        assert np.array_equal(det.anchor[b, :n].cpu().numpy(), anchors[b])
        assert np.array_equal(det.cls[b, :n].cpu().numpy(), rows[b][:, 5].astype(np.int32))
        assert np.array_equal(det.score[b, :n].cpu().numpy(), rows[b][:, 4])


def test_few_classes_large_segments_and_class_major():
    """nc = 2: ~1300 candidates per class -> CTA-wide class sort; nc = 1 with everything in one class."""
    for nc, seed in ((2, 5), (1, 6), (20, 7)):
        pred = torch.from_numpy(synth.yolov8_pred(seed, 3, 8400, nc=nc)).to(DEV)
        for order, md in ((ops.ORDER_CLASS_MAJOR, 0), (ops.ORDER_SCORE_DESC, 300)):
            _both(lambda: ops.pred_filter(pred, nc, 0.001), 0.6, ops.RULE_PER_CLASS, order, md, max_out=8400)


def test_fallback_when_image_does_not_fit_and_max_nms():
    """conf = 0: all 8400 anchors are candidates -> more positions than shared memory holds; max_nms binds."""
    pred = torch.from_numpy(synth.yolov8_pred(8, 2, 8400, nc=80)).to(DEV)
    _both(lambda: ops.pred_filter(pred, 80, 0.0), 0.7, ops.RULE_TORCHVISION_CPU, ops.ORDER_SCORE_DESC, 300)
    _both(lambda: ops.pred_filter(pred, 80, 0.001), 0.7, ops.RULE_TORCHVISION_CPU, ops.ORDER_SCORE_DESC, 300, max_nms=1500)
    _both(lambda: ops.pred_filter(pred, 80, 0.001), 0.7, ops.RULE_PER_CLASS, ops.ORDER_CLASS_MAJOR, 0, max_nms=700,
          max_out=8400)


def test_repeatable_on_the_same_candidates():
    """The fused call never modifies the candidate keys / counts, fallback images included."""
    pred = torch.from_numpy(synth.yolov8_pred(9, 2, 8400, nc=80)).to(DEV)
    for conf in (0.0, 0.001):
        c = ops.pred_filter(pred, 80, conf)
        k0, n0 = c.key.clone(), c.count.clone()
        a = ops.sort_nms(c, 0.7, max_det=300, max_nms=1000)
        b = ops.sort_nms(c, 0.7, max_det=300, max_nms=1000)
        assert torch.equal(c.count, n0) and torch.equal(a.count, b.count) and torch.equal(a.anchor, b.anchor)
        for i in range(2):
            assert torch.equal(c.key[i, :int(n0[i])], k0[i, :int(n0[i])])


def test_ssd_multi_key_per_prior_class_major():
    loc, conf = synth.ssd_head(13, 6)
    pri = torch.from_numpy(oracle.ssd_priors()).to(DEV)
    tl, tc = torch.from_numpy(loc).to(DEV), torch.from_numpy(conf).to(DEV)
    for thr in (0.001, 0.5):
        _both(lambda: ops.ssd_decode_filter(tl, tc, pri, thr, max_cand=32768), 0.5, ops.RULE_PER_CLASS,
              ops.ORDER_CLASS_MAJOR, 0, max_out=4096)
