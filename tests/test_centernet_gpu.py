"""-m gpu: CenterNet peak + top-K + DIoU-NMS kernels vs the oracle and the reference fixtures."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

import oracle
import synth
from gpu_util import record_error, BOX_ATOL, BOX_RTOL, SCORE_RTOL

pytestmark = pytest.mark.gpu

from computervision.pytorch_b200 import ops  # noqa: E402
from computervision.pytorch_b200.core.algorithms.centernet import CenterNetA  # noqa: E402
from computervision.pytorch_b200.core.utils.nms import diou_nms  # noqa: E402

DEV = "cuda:0"


def _cfg(nc, inp, use_nms=True):
    return NS(dataset=NS(num_classes=nc), arch=NS(input_size=(3,) + tuple(inp), downsampling_ratio=4),
              decode=NS(max_boxes_per_img=100, score_threshold=0.1, nms_threshold=0.5, use_nms=use_nms,
                        letterbox_image=True))


def _close_boxes(a, b, scale=1.0):
    record_error(a, b, "box")
    return a.shape == b.shape and np.all(np.abs(a - b) <= BOX_RTOL * np.abs(b) + BOX_ATOL * max(scale, 1.0))


def test_diou_nms_bit_exact_vs_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "centernet.npz"))
    for i in range(int(g["diou_cases"])):
        b, s = g[f"diou_boxes{i}"], g[f"diou_scores{i}"]
        keep = diou_nms(torch.from_numpy(b).to(DEV), torch.from_numpy(s).to(DEV), 0.5)
        assert keep.dtype == torch.int64 and not keep.is_cuda
        assert np.array_equal(keep.numpy(), g[f"diou_keep{i}"])


def test_decode_boxes_vs_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "centernet.npz"))
    for tag in g["cases"]:
        seed, B, H, W, nc, in_h, in_w = [int(v) for v in g[f"{tag}_cfg"]]
        pred = synth.centernet_pred(seed, B, H, W, nc)
        assert synth.checksum([pred]) == int(g[f"{tag}_crc"])
        t = torch.from_numpy(pred).to(DEV)
        for ntag, use_nms in (("nms", True), ("raw", False)):
            algo = CenterNetA(_cfg(nc, (in_h, in_w), use_nms), DEV)
            for ctag, conf in (("lo", 0.001), ("hi", 0.1)):
                for b in range(B):
                    boxes, scores, classes = algo.decode_boxes(t[b:b + 1], 480, 640, conf)
                    k = f"{tag}_{b}_{ntag}_{ctag}"
                    assert classes.dtype == np.int64 and np.array_equal(classes, g[k + "_classes"])
                    assert np.all(np.abs(scores - g[k + "_scores"]) <= SCORE_RTOL * g[k + "_scores"])
                    assert _close_boxes(boxes, g[k + "_boxes"], 640 / in_w)


@pytest.mark.parametrize("use_nms", [False, True])
def test_c3_batch64_vs_oracle(use_nms):
    """BASELINE config 3 at full size: (64, 128, 128, 84), K=100."""
    pred = synth.centernet_pred(4040, 64, 128, 128, 80)
    hw = [(480 + 7 * b, 640 - 3 * b) for b in range(64)]
    lb = ops.letterbox_params(hw, (512, 512), DEV)
    det = ops.centernet_decode(torch.from_numpy(pred).to(DEV), 100, 0.001, use_nms, 0.5, lb)
    ref = oracle.centernet_decode(pred, 100, 0.001, 0, use_nms, 0.5, oracle.letterbox_params(hw, (512, 512)))
    cnt = det.count.cpu().numpy()
    box, score, cls, pix = (t.cpu().numpy() for t in (det.box, det.score, det.cls, det.pixel))
    kept_lt_k = 0
    for b, (rb, rs, rc, rp) in enumerate(ref):
        n = int(cnt[b])
        assert n == len(rs)
        assert np.array_equal(cls[b, :n], rc) and np.array_equal(pix[b, :n], rp)
        record_error(score[b, :n], rs, "score")
        assert np.all(np.abs(score[b, :n] - rs) <= SCORE_RTOL * rs)
        assert _close_boxes(box[b, :n], rb, 2.0)
        kept_lt_k += n < 100
    if use_nms:
        assert kept_lt_k > 0          # DIoU-NMS really suppressed something


def test_stage_b_exact_and_edge_cases():
    # identical scores / plateau: ties resolve to the lower flat index; K larger than the number of peaks
    H = W = 8
    nc = 4
    pred = np.full((1, H, W, nc + 4), -20.0, np.float32)
    pred[..., nc:] = 0.5
    for (y, x, c) in [(1, 1, 0), (1, 5, 2), (6, 6, 3), (3, 3, 1)]:
        pred[0, y, x, c] = 2.0                      # four equal peaks
    det = ops.centernet_decode(torch.from_numpy(pred).to(DEV), 16, 0.5, False)
    ref = oracle.centernet_decode(pred, 16, 0.5)[0]
    n = int(det.count.item())
    assert n == len(ref[1]) == 4
    assert np.array_equal(det.pixel[0, :n].cpu().numpy(), ref[3]) and np.array_equal(det.cls[0, :n].cpu().numpy(), ref[2])
    assert np.array_equal(det.box[0, :n].cpu().numpy(), ref[0])      # no exp involved beyond the score: exact
    # nothing above the threshold
    det = ops.centernet_decode(torch.from_numpy(pred).to(DEV), 16, 0.95, True)
    assert int(det.count.item()) == 0
    # batch merge semantics of the reference-signature entry point
    pred2 = synth.centernet_pred(9, 2, 32, 32, 6, K=20)
    algo = CenterNetA(NS(dataset=NS(num_classes=6), arch=NS(input_size=(3, 128, 128), downsampling_ratio=4),
                         decode=NS(max_boxes_per_img=20, score_threshold=0.001, nms_threshold=0.5, use_nms=True,
                                   letterbox_image=True)), DEV)
    boxes, scores, classes = algo.decode_boxes(torch.from_numpy(pred2).to(DEV), 100, 150)
    r = oracle.centernet_decode(pred2, 20, 0.001, 0, False)
    mb = np.concatenate([x[0] for x in r])
    ms = np.concatenate([x[1] for x in r])
    mc = np.concatenate([x[2] for x in r])
    keep = oracle.diou_nms(mb, ms, 0.5)
    assert np.array_equal(classes, mc[keep].astype(np.int64))
    assert np.all(np.abs(scores - ms[keep]) <= SCORE_RTOL * ms[keep])


def test_list_overflow_is_redone_exactly():
    """Dense maps overflow the tile kernel's candidate list (no per-tile cap): the flagged images are redone by
    the exact row kernel.  A perfectly flat heatmap makes every cell a peak with the same score, so the top K are
    the K lowest flat indices (tie rule: lower index first); a small K on a random map overflows H*K as well."""
    H = W = 32
    nc = 8
    pred = torch.zeros((2, H, W, nc + 4), device=DEV)
    pred[..., :nc] = 0.3
    pred[..., nc:nc + 2] = 0.5
    pred[..., nc + 2:] = 2.0
    det = ops.centernet_decode(pred, 50, 0.001)
    assert det.count.tolist() == [50, 50]
    flat = (det.pixel.long() * nc + det.cls.long()).cpu()
    assert torch.equal(flat[0], torch.arange(50)) and torch.equal(flat[1], torch.arange(50))
    # random map, tiny K: H*K = 64 list slots, hundreds of peaks before a bound exists
    p2 = synth.centernet_pred(77, 2, 32, 32, 8)
    ref = oracle.centernet_decode(p2, 2, 0.001, 0, False, 0.5, None)
    got = ops.centernet_decode(torch.from_numpy(p2).to(DEV), 2, 0.001)
    for b, (box, score, cls, pix) in enumerate(ref):
        n = int(got.count[b])
        assert n == len(cls) and np.array_equal(got.cls[b, :n].cpu().numpy(), cls)
        assert np.array_equal(got.pixel[b, :n].cpu().numpy(), pix)


def test_dense_helpers_match_torch():
    """_suppress_redundant_centers (cvpp_centernet_suppress) and reverse_letter_box (cvpp_letterbox_reverse)
    against the reference's own eager arithmetic, bit for bit."""
    from computervision.pytorch_b200.core.utils.image_process import reverse_letter_box
    g = torch.Generator(device=DEV).manual_seed(3)
    heat = torch.rand((2, 9, 17, 7), generator=g, device=DEV)
    heat[0, 3, 5, 2] = heat[0, 3, 6, 3]                                   # a tie between neighbours: both are kept
    got = CenterNetA._suppress_redundant_centers(heat)
    hmax = torch.nn.functional.max_pool2d(heat, kernel_size=3, stride=1, padding=1)
    assert torch.equal(got, heat * (heat == hmax).float())
    boxes = torch.rand((37, 4), generator=g, device=DEV)
    for xywh in (True, False):
        for (h, w), inp in (((480, 640), (384, 384)), ((1080, 1920), (640, 640)), ((333, 500), (416, 416))):
            want = reverse_letter_box(h, w, list(inp), boxes.cpu(), xywh=xywh)   # the reference's eager CPU arithmetic
            assert torch.equal(reverse_letter_box(h, w, list(inp), boxes, xywh=xywh).cpu(), want)


@pytest.mark.parametrize("nc,H,W", [(6, 32, 32), (3, 32, 32), (10, 24, 40)])
def test_odd_class_counts_take_the_row_kernel(nc, H, W):
    """(nc + 4) % 4 != 0 (or tiny maps): columns are not 16-byte aligned, the exact row kernel does the whole map."""
    p = synth.centernet_pred(31 + nc, 2, H, W, nc)
    K = min(20, H * W * nc)
    ref = oracle.centernet_decode(p, K, 0.001, 0, False, 0.5, None)
    got = ops.centernet_decode(torch.from_numpy(p).to(DEV), K, 0.001)
    for b, (box, score, cls, pix) in enumerate(ref):
        n = int(got.count[b])
        assert n == len(cls) and np.array_equal(got.cls[b, :n].cpu().numpy(), cls)
        assert np.array_equal(got.pixel[b, :n].cpu().numpy(), pix)
