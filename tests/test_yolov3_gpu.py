"""-m gpu: YOLOv3 decode + per-(anchor, class) filter + per-class NMS vs the oracle and the reference
fixtures; dense predict_bounding_bbox; standalone yolo3_nms; gather_feat."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

import oracle
import synth
from gpu_util import record_error, BOX_RTOL, SCORE_RTOL, decode_keys

pytestmark = pytest.mark.gpu

from computervision.pytorch_b200 import ops  # noqa: E402
from computervision.pytorch_b200.core.loss.centernet_loss import RegL1Loss  # noqa: E402
from computervision.pytorch_b200.core.predict.yolov3_decode import Decoder, predict_bounding_bbox  # noqa: E402
from computervision.pytorch_b200.core.utils.anchor import generate_yolo3_anchor  # noqa: E402
from computervision.pytorch_b200.core.utils.nms import gather_op, yolo3_nms  # noqa: E402

DEV = "cuda:0"


def _cfg(nc=20):
    return NS(arch=NS(num_classes=nc, input_size=(3, 416, 416), anchor=list(oracle.YOLOV3_ANCHORS)),
              decode=NS(conf_threshold=0.6, iou_threshold=0.5))


def _close(got, ref, rtol, atol=0.0):
    record_error(got, ref, "box" if atol else "score_or_norm_box")
    return got.shape == ref.shape and bool(np.all(np.abs(got - ref) <= rtol * np.abs(ref) + atol))


def test_decoder_vs_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "yolov3.npz"))
    for tag in ("one", "two"):
        seed, B, merged = [int(v) for v in g[f"{tag}_cfg"]]
        levels = synth.yolov3_head(seed, B, merged=bool(merged))
        assert synth.checksum(levels) == int(g[f"{tag}_crc"])
        outs = [torch.from_numpy(l).to(DEV) for l in levels]
        for ctag, thr in (("eval", 0.001), ("pred", 0.6)):
            boxes, scores, classes = Decoder(_cfg(), thr, DEV)(outs)
            assert classes.dtype == torch.int32 and scores.dim() == 1
            assert np.array_equal(classes.cpu().numpy(), g[f"{tag}_{ctag}_classes"])
            assert _close(scores.cpu().numpy(), g[f"{tag}_{ctag}_scores"], SCORE_RTOL)
            assert _close(boxes.cpu().numpy(), g[f"{tag}_{ctag}_boxes"], BOX_RTOL, 1e-6)
        boxes, scores, classes = Decoder(_cfg(), 0.9999999, DEV)(outs)
        assert list(boxes.shape) + list(scores.shape) + list(classes.shape) == list(g[f"{tag}_empty_shapes"])


@pytest.mark.parametrize("B,nc,merged", [(4, 20, False), (3, 20, True), (2, 80, False)])
def test_kept_rows_exact_vs_oracle(B, nc, merged):
    levels = synth.yolov3_head(60 + B, B, nc=nc, merged=merged)
    outs = [torch.from_numpy(l).to(DEV) for l in levels]
    dec = Decoder(_cfg(nc), 0.001, DEV)
    if merged:
        boxes, scores = oracle.yolov3_dense(levels, nc)
        ob, os_, oc, orow, cands = oracle.yolo3_nms(boxes, scores, 0.001, 0.5)
        cand = dec._candidates(outs, merge_batch=True)
        assert int(cand.count.item()) == cands
        det = ops.per_class_nms_device(cand, 0.5)
        n = int(det.count.item())
        assert np.array_equal(det.anchor[0, :n].cpu().numpy(), orow)             # flattened row index: bit-exact
        assert np.array_equal(det.cls[0, :n].cpu().numpy(), oc)
        assert _close(det.score[0, :n].cpu().numpy(), os_, SCORE_RTOL) and _close(det.box[0, :n].cpu().numpy(), ob, BOX_RTOL, 1e-6)
        assert n < cands
        return
    det = dec.decode_batch(outs)
    cnt = det.count.cpu().numpy()
    for b in range(B):
        boxes, scores = oracle.yolov3_dense([l[b:b + 1] for l in levels], nc)
        ob, os_, oc, orow, cands = oracle.yolo3_nms(boxes, scores, 0.001, 0.5)
        assert int(det.cand_count[b].item()) == cands and cnt[b] == len(oc) and cnt[b] < cands
        assert np.array_equal(det.anchor[b, :cnt[b]].cpu().numpy(), orow)
        assert np.array_equal(det.cls[b, :cnt[b]].cpu().numpy(), oc)
        assert _close(det.score[b, :cnt[b]].cpu().numpy(), os_, SCORE_RTOL)
        assert _close(det.box[b, :cnt[b]].cpu().numpy(), ob, BOX_RTOL, 1e-6)


def test_generic_kernel_matches_stream_kernel(monkeypatch):
    levels = synth.yolov3_head(7, 2)
    ls = ops.make_levels([torch.from_numpy(l).to(DEV) for l in levels])
    anchors = np.array(oracle.YOLOV3_ANCHORS, np.float32).reshape(-1, 2)
    a = ops.yolov3_decode_filter(ls, 20, anchors, (416, 416), 0.001)
    monkeypatch.setenv("CVPP_FORCE_GENERIC", "1")
    b = ops.yolov3_decode_filter(ls, 20, anchors, (416, 416), 0.001)
    assert torch.equal(a.count, b.count)
    for i in range(2):
        n = int(a.count[i])
        ka, kb = a.key[i, :n].sort().values, b.key[i, :n].sort().values
        assert torch.equal(ka, kb)
        anc = (ka & 0x1FFFFF).long()
        assert torch.equal(a.box_dense[i, anc], b.box_dense[i, anc])


def test_predict_bounding_bbox_dense(golden_dir):
    g = np.load(os.path.join(golden_dir, "yolov3.npz"))
    levels = synth.yolov3_head(51, 1)
    cfg = _cfg()
    f = torch.from_numpy(levels[0]).to(DEV)
    xy, wh, conf, prob = predict_bounding_bbox(20, f, generate_yolo3_anchor(cfg, DEV, 0), DEV)
    assert _close(xy.cpu().numpy(), g["pbb_xy"], 1e-5, 1e-7) and _close(wh.cpu().numpy(), g["pbb_wh"], 1e-5)
    assert _close(conf.cpu().numpy(), g["pbb_conf"], 1e-5)
    assert _close(prob.cpu().numpy().reshape(-1)[::7], g["pbb_prob_sub"], 1e-5)
    xy2, wh2, grid, fm = predict_bounding_bbox(20, f, generate_yolo3_anchor(cfg, DEV, 0), DEV, is_training=True)
    assert torch.equal(xy2, xy) and np.array_equal(grid.cpu().numpy(), g["pbb_grid"])
    assert np.array_equal(fm.cpu().numpy().reshape(-1)[::11], g["pbb_fm_sub"])
    # a batch, every scale, against the oracle (flattened order)
    levels = synth.yolov3_head(8, 3)
    for i, l in enumerate(levels):
        xy, wh, conf, prob = predict_bounding_bbox(20, torch.from_numpy(l).to(DEV), generate_yolo3_anchor(cfg, DEV, i), DEV)
        boxes, scores = oracle.yolov3_dense([l], 20, anchors=oracle.YOLOV3_ANCHORS[6 * i:6 * i + 6])
        got = torch.cat((xy - wh / 2, xy + wh / 2), -1).reshape(-1, 4).cpu().numpy()
        assert _close(got, boxes, BOX_RTOL, 1e-6)
        assert _close((conf * prob).reshape(-1, 20).cpu().numpy(), scores, 2e-5)


def test_standalone_yolo3_nms_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "yolov3.npz"))
    b, s = torch.from_numpy(g["nms3_boxes_in"]).to(DEV), torch.from_numpy(g["nms3_scores_in"]).to(DEV)
    ob, os_, oc = yolo3_nms(5, 0.5, 0.45, b, s, DEV)
    assert np.array_equal(ob.cpu().numpy(), g["nms3_boxes"]) and np.array_equal(os_.cpu().numpy(), g["nms3_scores"])
    assert np.array_equal(oc.cpu().numpy(), g["nms3_classes"]) and oc.dtype == torch.int32
    ob, os_, oc = yolo3_nms(5, 1.0, 0.45, b, s, DEV)
    assert ob.shape == (0, 4) and os_.shape == (0, 1) and oc.shape == (0,)


def test_gather_feat_and_gather_op():
    rng = synth.rng_for(1)
    feat = torch.from_numpy(rng.standard_normal((3, 16, 12, 2), dtype=np.float32)).to(DEV)
    ind = torch.from_numpy(rng.integers(0, 16 * 12, (3, 40)).astype(np.int32)).to(DEV)
    got = RegL1Loss.gather_feat(feat, ind)
    want = torch.gather(feat.reshape(3, -1, 2), 1, ind.long().unsqueeze(2).expand(-1, -1, 2))
    assert torch.equal(got, want)
    assert torch.equal(ops.gather_feat(feat.reshape(3, -1, 2), ind.long()), want)
    with pytest.raises(IndexError):
        ops.gather_feat(feat.reshape(3, -1, 2), ind + 1000)
    t = torch.from_numpy(rng.standard_normal((50, 4), dtype=np.float32)).to(DEV)
    idx = torch.tensor([3, 3, 49, 0], device=DEV)
    assert torch.equal(gather_op(t, idx, DEV), t[idx])
    assert torch.equal(gather_op(t[:, 0].contiguous(), idx, DEV), t[idx, :1])
