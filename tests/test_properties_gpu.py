"""-m gpu: property tests (hypothesis) of the CUDA sort + NMS through the C ABI - SURVEY.md §4's list:
permutation invariance (tie-free scores), class-awareness (different classes never suppress each other),
idempotence, subset / order, agreement of the fused and the two-kernel path, and agreement with the oracle on
arbitrary (also degenerate) boxes."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import oracle

pytestmark = pytest.mark.gpu

from computervision.pytorch_b200 import ops  # noqa: E402
from test_nms_golden_gpu import gpu_nms_both  # noqa: E402

SET = dict(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
RULES = (ops.RULE_TORCHVISION_CPU, ops.RULE_COORD_TRICK, ops.RULE_PER_CLASS)


def _case(seed, n, nc, quantised):
    rng = np.random.Generator(np.random.PCG64(seed))
    k = max(n // 6, 1)
    ctr = rng.uniform(50, 590, (k, 2))[rng.integers(0, k, n)] + rng.normal(0, 6, (n, 2))
    wh = np.exp(rng.normal(3.5, 0.5, (n, 2)))
    boxes = np.concatenate([ctr - wh / 2, ctr + wh / 2], 1)
    if quantised:
        boxes = np.round(boxes / 8) * 8          # many exactly equal coordinates and rational IoUs
    scores = ((rng.permutation(n) + 1) / np.float32(n + 1)).astype(np.float32)   # tie-free
    cls = rng.integers(0, nc, n)
    return boxes.astype(np.float32), scores, cls


@settings(**SET)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(1, 1400), nc=st.integers(1, 12), q=st.booleans(),
       thr=st.sampled_from([0.3, 0.45, 0.5, 0.7]), rule=st.sampled_from(RULES))
def test_matches_oracle_and_is_permutation_invariant(seed, n, nc, q, thr, rule):
    boxes, scores, cls = _case(seed, n, nc, q)
    want = oracle.batched_nms(boxes, scores, cls, thr, {ops.RULE_TORCHVISION_CPU: 0, ops.RULE_COORD_TRICK: 1,
                                                        ops.RULE_PER_CLASS: 2}[rule])
    a, z = gpu_nms_both(boxes, scores, thr, cls=cls, nc=nc, rule=rule)
    assert np.array_equal(a, want) and np.array_equal(z, want)
    # subset + strictly descending scores
    assert len(set(a.tolist())) == len(a) and np.all(np.diff(scores[a]) < 0)
    # permutation invariance: shuffle the input rows, the same ORIGINAL boxes survive in the same order
    perm = np.random.Generator(np.random.PCG64(seed ^ 0x5bd1e995)).permutation(n)
    ap, zp = gpu_nms_both(boxes[perm], scores[perm], thr, cls=cls[perm], nc=nc, rule=rule)
    assert np.array_equal(perm[ap], a) and np.array_equal(perm[zp], a)
    # idempotence: NMS of the survivors keeps all of them
    ai, zi = gpu_nms_both(boxes[a], scores[a], thr, cls=cls[a], nc=nc, rule=rule)
    if rule != ops.RULE_TORCHVISION_CPU or (len(a) <= 1000) == (n <= 1000):   # same batched_nms branch both times
        assert np.array_equal(ai, np.arange(len(a))) and np.array_equal(zi, np.arange(len(a)))


@settings(**SET)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(2, 300), nc=st.integers(2, 40), rule=st.sampled_from(RULES))
def test_different_classes_never_suppress_each_other(seed, n, nc, rule):
    """n copies of ONE box: exactly the best-scored copy of every class survives, whatever the threshold."""
    rng = np.random.Generator(np.random.PCG64(seed))
    box = np.array([10.0, 20.0, 200.0, 300.0], np.float32)
    boxes = np.repeat(box[None], n, 0)
    scores = ((rng.permutation(n) + 1) / np.float32(n + 1)).astype(np.float32)
    cls = rng.integers(0, nc, n)
    a, z = gpu_nms_both(boxes, scores, 0.5, cls=cls, nc=nc, rule=rule)
    best = np.array(sorted((max((i for i in range(n) if cls[i] == c), key=lambda i: scores[i])
                            for c in set(cls.tolist())), key=lambda i: -scores[i]), np.int64)
    assert np.array_equal(a, best) and np.array_equal(z, best)


@settings(**SET)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(1, 400))
def test_degenerate_boxes_match_oracle(seed, n):
    """zero-area, inverted and repeated boxes (IoU 0/0 = NaN never suppresses) - same answers as the oracle."""
    rng = np.random.Generator(np.random.PCG64(seed))
    boxes = rng.integers(0, 6, (n, 4)).astype(np.float32)        # x2 < x1 and w = 0 happen all the time
    scores = ((rng.permutation(n) + 1) / np.float32(n + 1)).astype(np.float32)
    for thr in (0.0, 0.5):
        want = oracle.nms(boxes, scores, thr)
        a, z = gpu_nms_both(boxes, scores, thr)
        assert np.array_equal(a, want) and np.array_equal(z, want)
