"""-m gpu: cvpp_topk (CenterNetA._top_k, reference core/algorithms/centernet.py:328-338) - exact radix select with
the product's tie rule (equal scores ordered by the lower flat index), checked against the reference fixture
(tests/golden/centernet_topk.npz, written by the real _suppress_redundant_centers + _top_k), the oracle on
tie-heavy maps, and torch.topk's values."""
import os

import numpy as np
import pytest
import torch

import oracle
import synth

pytestmark = pytest.mark.gpu

from computervision.pytorch_b200 import ops  # noqa: E402
from computervision.pytorch_b200.core.algorithms.centernet import CenterNetA  # noqa: E402

DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_mirror_top_k_matches_reference_fixture():
    g = np.load(os.path.join(GOLD, "centernet_topk.npz"))
    seed, B, H, W, nc = (int(v) for v in g["cfg"])
    pred = synth.centernet_pred(seed, B, H, W, nc)
    assert synth.checksum([pred]) == int(g["crc"]), "regenerated input drifted from the fixture"
    heat = CenterNetA._suppress_redundant_centers(torch.sigmoid(torch.from_numpy(pred[..., :nc]).to(DEV)))
    # sigmoid on the GPU differs from the CPU's by an ulp here and there; the selection must not
    assert np.allclose(heat.cpu().numpy().reshape(-1)[::97], g["heat_sub"], rtol=1e-5, atol=0)
    for K in (100, 7):
        sc, inds, cls, ys, xs = CenterNetA._top_k(heat, K)
        assert inds.dtype == torch.int32 and cls.dtype == torch.int64
        assert np.array_equal(inds.cpu().numpy(), g[f"k{K}_inds"])
        assert np.array_equal(cls.cpu().numpy(), g[f"k{K}_cls"])
        assert np.array_equal(ys.cpu().numpy(), g[f"k{K}_ys"]) and np.array_equal(xs.cpu().numpy(), g[f"k{K}_xs"])
        assert np.allclose(sc.cpu().numpy(), g[f"k{K}_scores"], rtol=1e-5, atol=0)


@pytest.mark.parametrize("B,N,K", [(3, 128 * 128 * 80, 100), (2, 5000, 4096), (5, 33, 33), (1, 1, 1), (4, 100003, 257)])
def test_tie_heavy_rows_lower_index_first(B, N, K):
    """Rows made of a handful of distinct values (and long runs of exact zeros, like a suppressed heat map): the
    boundary group is far larger than the sort capacity, so the select descends into the index digits."""
    rng = np.random.Generator(np.random.PCG64(N + K))
    s = np.zeros((B, N), np.float32)
    for b in range(B):
        vals = np.array([0.0, 0.0, 0.0, 0.25, 0.5, 0.5, 0.75, -1.0, 1e-30], np.float32)
        s[b] = vals[rng.integers(0, len(vals), N)]
        if b == 1:
            s[b] = 0.0                                   # everything ties
        if b == 2:
            s[b, rng.integers(0, N, 5)] = np.float32(0.9)  # a few clear winners, then ties
    want_v, want_i = oracle.topk(s, K)
    v, i = ops.topk(torch.from_numpy(s).to(DEV), K)
    assert np.array_equal(i.cpu().numpy(), want_i)
    assert np.array_equal(v.cpu().numpy(), want_v)


@pytest.mark.parametrize("misalign", [0, 1])
def test_random_rows_match_torch_and_oracle(misalign):
    g = torch.Generator(device=DEV)
    g.manual_seed(5)
    for B, N, K in ((4, 70001, 100), (2, 96 * 96 * 20, 1000), (64, 4096, 16)):
        buf = torch.randn((B * N + 1,), generator=g, device=DEV)
        s = buf[misalign:misalign + B * N].view(B, N)       # rows that are not 16-byte aligned take the scalar loads
        v, i = ops.topk(s, K, split=None)
        tv, _ = torch.topk(s, K, dim=1, largest=True, sorted=True)
        assert torch.equal(v, tv)
        ov, oi = oracle.topk(s.cpu().numpy(), K)
        assert np.array_equal(i.cpu().numpy(), oi) and np.array_equal(v.cpu().numpy(), ov)


def test_split_outputs_and_errors():
    s = torch.rand((2, 6 * 5 * 7), device=DEV)
    v, i, c, y, x, pix = ops.topk(s, 9, split=(7, 5))
    assert torch.equal(c, i % 7) and torch.equal(y, (i // 7) // 5) and torch.equal(x, (i // 7) % 5)
    assert pix.dtype == torch.int32 and torch.equal(pix.long(), y * 5 + x)
    with pytest.raises(RuntimeError):
        ops.topk(s, 6 * 5 * 7 + 1)
    with pytest.raises(ValueError):
        ops.topk(s.cpu(), 3)
    nan = torch.tensor([[0.5, float("nan"), 2.0, -1.0]], device=DEV)
    v, i = ops.topk(nan, 2)
    assert i.tolist() == [[1, 2]] and torch.isnan(v[0, 0])    # NaN first, like torch.topk
