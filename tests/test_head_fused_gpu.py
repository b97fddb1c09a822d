"""-m gpu: cvpp_yolov8_head_decode_filter - the head's last 1x1 convolutions (reference
core/models/yolov8/modules.py:423-425,431) fused with the decode + filter on the tensor cores (tcgen05 kind::tf32).

Parity design: the GEMM's inputs are chosen so that the result does not depend on the arithmetic of the contraction,
then everything downstream must be BIT-identical to the streaming decode kernel on the materialised head:
  * dyadic test: activations k/8, weights k/64, biases k/8 - every operand is exact in TF32, every product and every
    partial sum is exact in fp32 (checked against a float64 contraction), so the fused kernel must emit exactly the
    candidate keys and boxes cvpp_yolov8_decode_filter emits on conv(x) computed on the host;
  * random test: N(0,1) activations - the tensor core reads fp32 as TF32; the kernel is compared with the decode of
    a float64 contraction of TF32-truncated operands, within the tolerance of fp32 accumulation order."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from computervision.pytorch_b200 import ops  # noqa: E402
from computervision.pytorch_b200._lib import CvppError  # noqa: E402

DEV = "cuda:0"
STRIDES = (8.0, 16.0, 32.0)


def _head64(feats, ws, bs):
    """per level: (B, K, H, W) x (N, K) + (N) -> (B, N, H, W) in float64"""
    return [np.einsum("bkhw,nk->bnhw", f.astype(np.float64), w.astype(np.float64)) + b.astype(np.float64)[None, :, None, None]
            for f, w, b in zip(feats, ws, bs)]


def _dyadic_case(seed, B, sizes, c2, c3, nc, cls_bias=None):
    rng = np.random.Generator(np.random.PCG64(seed))
    if cls_bias is None:
        # class logits = sum of c3 products of U[-4,4] x U[-.5,.5] (sigma^2 ~ 0.467 c3): put the conf=.001 cut (logit -6.9)
        # ~2.6 sigma above the mean, so that a cell has a class above it with probability ~0.3
        cls_bias = -np.round((6.9 + 2.62 * np.sqrt(0.467 * c3)) * 8) / 8
    bf = [(rng.integers(-32, 33, (B, c2, h, w)) / 8.0).astype(np.float32) for h, w in sizes]
    cf = [(rng.integers(-32, 33, (B, c3, h, w)) / 8.0).astype(np.float32) for h, w in sizes]
    bw = [(rng.integers(-32, 33, (64, c2)) / 64.0).astype(np.float32) for _ in sizes]
    cw = [(rng.integers(-32, 33, (nc, c3)) / 64.0).astype(np.float32) for _ in sizes]
    bb = [(rng.integers(-16, 17, (64,)) / 8.0).astype(np.float32) for _ in sizes]
    cb = [(cls_bias + rng.integers(-16, 17, (nc,)) / 8.0).astype(np.float32) for _ in sizes]
    return bf, cf, bw, cw, bb, cb


def _materialised_head(bf, cf, bw, cw, bb, cb, exact=True):
    hb, hc = _head64(bf, bw, bb), _head64(cf, cw, cb)
    head = [np.concatenate([a, c], 1) for a, c in zip(hb, hc)]
    head32 = [h.astype(np.float32) for h in head]
    if exact:
        for h, h32 in zip(head, head32):
            assert np.array_equal(h, h32.astype(np.float64)), "the test data must make the contraction exact in fp32"
    return head32


def _to(ts):
    return [torch.from_numpy(np.ascontiguousarray(t)).to(DEV) for t in ts]


def _sorted_keys(c):
    key = c.key.cpu().numpy().view(np.uint64)
    cnt = c.count.cpu().numpy()
    return [np.sort(key[b, :cnt[b]]) for b in range(key.shape[0])], cnt


@pytest.mark.parametrize("B,sizes,c2,c3,nc", [
    (3, ((80, 80), (40, 40), (20, 20)), 64, 80, 80),      # the n model of the reference, 640 x 640
    (1, ((20, 20),), 64, 80, 80),                          # bs = 1, one level with a partial last tile (400 cells)
    (2, ((40, 40), (20, 20)), 64, 128, 20),                # s-model width, VOC classes (nc padded to 32 columns)
    (5, ((8, 12), (4, 6)), 80, 320, 80),                   # x-model widths, tiny odd grids (96 and 24 cells)
])
def test_fused_head_is_bit_identical_to_conv_plus_decode(B, sizes, c2, c3, nc):
    bf, cf, bw, cw, bb, cb = _dyadic_case(7 + B, B, sizes, c2, c3, nc)
    head = _materialised_head(bf, cf, bw, cw, bb, cb)
    strides = STRIDES[:len(sizes)]
    want = ops.yolov8_decode_filter(ops.make_levels(_to(head), strides), nc, 0.001)
    got = ops.yolov8_head_decode_filter(_to(bf), _to(cf), _to(bw), _to(bb), _to(cw), _to(cb), strides, 0.001)
    torch.cuda.synchronize()
    wk, wc = _sorted_keys(want)
    gk, gc = _sorted_keys(got)
    assert np.array_equal(gc, wc), (gc, wc)
    assert wc.sum() > 3 * B and (wc < 0.8 * got.A).all(), "the case must be selective, not empty and not everything"
    wbox, gbox = want.box_dense.cpu().numpy(), got.box_dense.cpu().numpy()
    for b in range(B):
        assert np.array_equal(gk[b], wk[b]), b
        anchors = (wk[b] & np.uint64(0x1FFFFF)).astype(np.int64)
        assert np.array_equal(gbox[b, anchors], wbox[b, anchors])


def test_fused_head_then_nms_equals_the_unfused_pipeline():
    B, sizes = 4, ((80, 80), (40, 40), (20, 20))
    bf, cf, bw, cw, bb, cb = _dyadic_case(99, B, sizes, 64, 80, 80)
    head = _materialised_head(bf, cf, bw, cw, bb, cb)
    ref = ops.sort_nms(ops.yolov8_decode_filter(ops.make_levels(_to(head), STRIDES), 80, 0.001), 0.7, max_det=300, max_nms=30000)
    got = ops.sort_nms(ops.yolov8_head_decode_filter(_to(bf), _to(cf), _to(bw), _to(bb), _to(cw), _to(cb), STRIDES, 0.001),
                       0.7, max_det=300, max_nms=30000)
    torch.cuda.synchronize()
    assert torch.equal(got.count, ref.count) and int(ref.count.min()) > 0
    for b, n in enumerate(ref.count.tolist()):
        assert torch.equal(got.anchor[b, :n], ref.anchor[b, :n]) and torch.equal(got.cls[b, :n], ref.cls[b, :n])
        assert torch.equal(got.score[b, :n], ref.score[b, :n]) and torch.equal(got.box[b, :n], ref.box[b, :n])


def test_full_c2_shape_repeated_launches_are_identical_and_exact():
    """BASELINE configs[1] shape (64 images x 8400 cells): every CTA walks ~29 tiles, the ring wraps ~18 times and both
    accumulator stages are reused ~14 times per launch.  40 back-to-back launches (the pipeline state of one must not leak
    into the next) must all equal the decode of an exact float64 contraction."""
    B, sizes = 64, ((80, 80), (40, 40), (20, 20))
    g = torch.Generator(device=DEV)
    g.manual_seed(2024)
    ri = lambda shape, lo, hi, div: (torch.randint(lo, hi, shape, generator=g, device=DEV).float() / div)   # noqa: E731
    bf = [ri((B, 64, h, w), -32, 33, 8.0) for h, w in sizes]
    cf = [ri((B, 80, h, w), -32, 33, 8.0) for h, w in sizes]
    bw = [ri((64, 64), -32, 33, 64.0) for _ in sizes]
    cw = [ri((80, 80), -32, 33, 64.0) for _ in sizes]
    bb = [ri((64,), -16, 17, 8.0) for _ in sizes]
    cls_bias = -round((6.9 + 2.62 * (0.467 * 80) ** 0.5) * 8) / 8
    cb = [ri((80,), -16, 17, 8.0) + cls_bias for _ in sizes]
    head = []
    for l in range(3):
        hb = torch.einsum("bkhw,nk->bnhw", bf[l].double(), bw[l].double()) + bb[l].double()[None, :, None, None]
        hc = torch.einsum("bkhw,nk->bnhw", cf[l].double(), cw[l].double()) + cb[l].double()[None, :, None, None]
        h64 = torch.cat((hb, hc), 1)
        h32 = h64.float()
        assert torch.equal(h32.double(), h64)
        head.append(h32)
        del hb, hc, h64
    want = ops.yolov8_decode_filter(ops.make_levels(head, STRIDES), 80, 0.001)
    wkey = torch.sort(torch.where(torch.arange(want.key.shape[1], device=DEV)[None] < want.count[:, None], want.key,
                                  torch.full_like(want.key, torch.iinfo(torch.int64).max)), dim=1).values
    assert int(want.count.min()) > 500
    mask = torch.zeros((B, want.A), dtype=torch.bool, device=DEV)
    for _ in range(40):
        got = ops.yolov8_head_decode_filter(bf, cf, bw, bb, cw, cb, STRIDES, 0.001)
        assert torch.equal(got.count, want.count)
        gkey = torch.sort(torch.where(torch.arange(got.key.shape[1], device=DEV)[None] < got.count[:, None], got.key,
                                      torch.full_like(got.key, torch.iinfo(torch.int64).max)), dim=1).values
        assert torch.equal(gkey, wkey)
        if not mask.any():
            valid = torch.arange(got.key.shape[1], device=DEV)[None] < got.count[:, None]
            anchors = (got.key & 0x1FFFFF).long()
            mask.scatter_(1, torch.where(valid, anchors, anchors[:, :1]), True)
        assert torch.equal(got.box_dense[mask], want.box_dense[mask])


def _tf32_trunc(a):
    return (a.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def test_random_inputs_match_a_tf32_contraction_within_accumulation_tolerance():
    rng = np.random.Generator(np.random.PCG64(5))
    B, sizes, c2, c3, nc = 2, ((40, 40), (20, 20)), 64, 80, 80
    bf = [rng.standard_normal((B, c2, h, w), dtype=np.float32) for h, w in sizes]
    cf = [rng.standard_normal((B, c3, h, w), dtype=np.float32) for h, w in sizes]
    bw = [(rng.standard_normal((64, c2), dtype=np.float32) * 0.4) for _ in sizes]
    cw = [(rng.standard_normal((nc, c3), dtype=np.float32) * 0.5) for _ in sizes]
    bb = [rng.standard_normal((64,), dtype=np.float32) for _ in sizes]
    cb = [(rng.standard_normal((nc,), dtype=np.float32) - 16.0) for _ in sizes]
    strides = STRIDES[1:]
    got = ops.yolov8_head_decode_filter(_to(bf), _to(cf), _to(bw), _to(bb), _to(cw), _to(cb), strides, 0.001)
    torch.cuda.synchronize()
    gk, gc = _sorted_keys(got)
    t = lambda xs: [_tf32_trunc(np.ascontiguousarray(x)) for x in xs]   # noqa: E731
    head = _materialised_head(t(bf), t(cf), t(bw), t(cw), bb, cb, exact=False)
    want = ops.yolov8_decode_filter(ops.make_levels(_to(head), strides), nc, 0.001)
    wk, wc = _sorted_keys(want)
    assert wc.sum() > 100
    gbox, wbox = got.box_dense.cpu().numpy(), want.box_dense.cpu().numpy()
    for b in range(B):
        def unpack(k):
            anchor = (k & np.uint64(0x1FFFFF)).astype(np.int64)
            score = (np.uint32(0x7FFFFFFF) - ((k >> np.uint64(21)) & np.uint64(0x7FFFFFFF)).astype(np.uint32)).view(np.float32)
            return dict(zip(anchor.tolist(), zip((k >> np.uint64(52)).astype(np.int64).tolist(), score.tolist())))
        g, w = unpack(gk[b]), unpack(wk[b])
        # anchors whose score is not within 1e-3 of the threshold are candidates on both sides
        for a, (c, s) in w.items():
            if s > 0.001 * 1.001:
                assert a in g, (b, a, s)
        common = sorted(set(g) & set(w))
        assert len(common) >= 0.98 * len(w)
        same_cls = sum(int(g[a][0] == w[a][0]) for a in common)
        assert same_cls >= 0.995 * len(common)          # a class flips only when two logits are within ~1e-5
        gs, ws = np.array([g[a][1] for a in common]), np.array([w[a][1] for a in common])
        assert np.all(np.abs(gs - ws) <= 2e-4 * ws + 1e-7), float(np.abs(gs - ws).max())
        assert np.all(np.abs(gbox[b, common] - wbox[b, common]) <= 1e-4 * np.abs(wbox[b, common]) + 5e-3)


def test_bad_shapes_are_rejected():
    bf, cf, bw, cw, bb, cb = _dyadic_case(1, 1, ((20, 20),), 64, 80, 80)
    with pytest.raises((CvppError, ValueError)):
        ops.yolov8_head_decode_filter(_to([bf[0][:, :40]]), _to(cf), _to([bw[0][:, :40]]), _to(bb), _to(cw), _to(cb), (32.0,), 0.001)
    with pytest.raises(ValueError):
        ops.yolov8_head_decode_filter(_to(bf), _to(cf), _to(bw), _to(bb), _to([cw[0][:, :64]]), _to(cb), (32.0,), 0.001)
    with pytest.raises(ValueError):
        ops.yolov8_head_decode_filter([torch.from_numpy(bf[0])], _to(cf), _to(bw), _to(bb), _to(cw), _to(cb), (32.0,), 0.001)
