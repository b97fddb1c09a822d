"""-m gpu: SSD prior decode + per-class NMS vs the oracle and the reference fixture."""
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
from computervision.pytorch_b200.core.algorithms.ssd import Ssd  # noqa: E402

DEV = "cuda:0"
AR = [[1, 2, 0.5], [1, 2, 0.5, 3, 1.0 / 3], [1, 2, 0.5, 3, 1.0 / 3], [1, 2, 0.5, 3, 1.0 / 3], [1, 2, 0.5], [1, 2, 0.5]]


def _cfg():
    return NS(arch=NS(input_size=(3, 300, 300), anchor_sizes=[30, 60, 111, 162, 213, 264, 315],
                      feature_shapes=[38, 19, 10, 5, 3, 1], aspect_ratios=AR),
              dataset=NS(num_classes=20), loss=NS(variance=[0.1, 0.2]),
              decode=NS(letterbox_image=True, conf_threshold=0.7, nms_threshold=0.5))


def test_decode_boxes_vs_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "ssd.npz"))
    loc, conf = synth.ssd_head(int(g["seed"][0]), int(g["seed"][1]))
    assert synth.checksum([loc, conf]) == int(g["crc"])
    algo = Ssd(_cfg(), DEV)
    tl, tc = torch.from_numpy(loc).to(DEV), torch.from_numpy(conf).to(DEV)
    sub = algo._parse_mbox_loc(tl[0]).cpu().numpy()[::37]
    assert np.all(np.abs(sub - g["parse_sub"]) <= 1e-5 * np.abs(g["parse_sub"]) + 1e-6)
    for ctag, thr in (("eval", 0.001), ("pred", None)):
        res = algo.decode_boxes((tl, tc), 480, 640, thr)
        for b, r in enumerate(res):
            ref = g[f"rows_{b}_{ctag}"]
            assert r.dtype == np.float32 and r.shape == ref.shape
            assert np.array_equal(r[:, 4], ref[:, 4])                       # labels, class-major order
            record_error(r[:, 5], ref[:, 5], "score")
            record_error(r[:, :4], ref[:, :4], "box")
            assert np.all(np.abs(r[:, 5] - ref[:, 5]) <= SCORE_RTOL * ref[:, 5])
            assert np.all(np.abs(r[:, :4] - ref[:, :4]) <= BOX_RTOL * np.abs(ref[:, :4]) + BOX_ATOL * 2.2)
    res = algo.decode_boxes((tl, tc), 480, 640, 0.99999)
    assert all(isinstance(r, list) and len(r) == 0 for r in res)


def test_c4_batch128_vs_oracle_prior_indices_exact():
    """BASELINE config 4 at full size: 128 x 8732 priors x 21 classes, eval threshold."""
    loc, conf = synth.ssd_head(77, 128)
    pri = oracle.ssd_priors()
    ref = oracle.ssd_decode(loc, conf, pri, 0.001, 0.5)
    cand = ops.ssd_decode_filter(torch.from_numpy(loc).to(DEV), torch.from_numpy(conf).to(DEV),
                                 torch.from_numpy(pri).to(DEV), 0.001, max_cand=32768)
    rows = ops.per_class_nms_rows(cand, 0.5)
    suppressed = 0
    for b, ((rrows, rprior), (box, score, cls, anchor)) in enumerate(zip(ref, rows)):
        assert np.array_equal(anchor.numpy(), rprior), b                    # kept prior indices: bit-exact
        assert np.array_equal(cls.numpy().astype(np.float32), rrows[:, 4])
        assert np.all(np.abs(score.numpy() - rrows[:, 5]) <= SCORE_RTOL * rrows[:, 5])
        assert np.all(np.abs(box.numpy() - rrows[:, :4]) <= 1e-5 * np.abs(rrows[:, :4]) + 1e-6)
        suppressed += int(cand.count[b].item()) - len(rprior)
    assert suppressed > 0


def test_stage_b_exact_on_identical_inputs():
    """Per-class NMS on the GPU's own decoded boxes / scores equals the oracle's nms loop bit for bit."""
    loc, conf = synth.ssd_head(3, 4)
    pri = oracle.ssd_priors()
    cand = ops.ssd_decode_filter(torch.from_numpy(loc).to(DEV), torch.from_numpy(conf).to(DEV),
                                 torch.from_numpy(pri).to(DEV), 0.01)
    key = cand.key.cpu().numpy().view(np.uint64)
    cnt = cand.count.cpu().numpy()
    dense = cand.box_dense.cpu().numpy()
    rows = ops.per_class_nms_rows(cand, 0.5)
    from gpu_util import decode_keys
    for b in range(4):
        cls, score, prior = decode_keys(key[b, :cnt[b]])
        o = np.lexsort((prior, cls))                                       # candidate order: class, then prior
        keep = oracle.nms_per_class(dense[b, prior[o]], score[o], cls[o].astype(np.int32), 20, 0.5)
        box, s, c, a = rows[b]
        assert np.array_equal(a.numpy(), prior[o][keep]) and np.array_equal(c.numpy(), cls[o][keep])
        assert np.array_equal(s.numpy(), score[o][keep]) and np.array_equal(box.numpy(), dense[b, prior[o][keep]])


@pytest.mark.parametrize("B,P,nc", [(3, 1000, 3),    # nc + 1 = 4: streaming kernel, generic class count, ragged last tile
                                    (2, 1001, 20),   # P (nc + 1) not a multiple of 4: block kernel (misaligned rows)
                                    (5, 36, 20),     # fewer tiles than SMs
                                    (1, 4100, 6)])
def test_ragged_shapes_vs_oracle(B, P, nc):
    """Shapes off the VOC fast path: both kernels against the oracle, kept prior indices bit-exact."""
    rng = np.random.default_rng(100 * B + P + nc)
    loc = rng.standard_normal((B, P, 4)).astype(np.float32)
    conf = (rng.standard_normal((B, P, nc + 1)) * 2.0).astype(np.float32)
    conf[..., 0] += 3.0
    # score separation as in synth.ssd_head: keep every probability clear of the threshold and of its neighbours
    cx, cy = rng.random((P,)) * 0.8 + 0.1, rng.random((P,)) * 0.8 + 0.1
    w, h = rng.random((P,)) * 0.15 + 0.02, rng.random((P,)) * 0.15 + 0.02
    pri = np.stack([cx - w, cy - h, cx + w, cy + h], 1).astype(np.float32)
    # threshold in the middle of the widest gap of the probabilities around 0.02: no score sits on the boundary
    e = np.exp(conf.astype(np.float64) - conf.max(-1, keepdims=True))
    prob = np.sort((e / e.sum(-1, keepdims=True))[..., 1:].ravel())
    near = prob[(prob > 0.015) & (prob < 0.025)]
    k = int(np.argmax(np.diff(near)))
    thr = float(np.float32((near[k] + near[k + 1]) / 2))
    assert near[k + 1] - near[k] > 4e-6
    ref = oracle.ssd_decode(loc, conf, pri, thr, 0.45)
    cand = ops.ssd_decode_filter(torch.from_numpy(loc).to(DEV), torch.from_numpy(conf).to(DEV),
                                 torch.from_numpy(pri).to(DEV), thr, max_cand=P * nc)
    rows = ops.per_class_nms_rows(cand, 0.45)
    total = 0
    for b, ((rrows, rprior), (box, score, cls, anchor)) in enumerate(zip(ref, rows)):
        assert np.array_equal(anchor.numpy(), rprior), b
        assert np.array_equal(cls.numpy().astype(np.float32), rrows[:, 4])
        assert np.all(np.abs(score.numpy() - rrows[:, 5]) <= SCORE_RTOL * rrows[:, 5])
        assert np.all(np.abs(box.numpy() - rrows[:, :4]) <= 1e-5 * np.abs(rrows[:, :4]) + 1e-6)
        total += len(rprior)
    assert total > 0
