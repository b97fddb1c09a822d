"""VOC-style mAP: the oracle's restatement of get_map's matching loop + the product's host
AP integration against the fixture the REAL get_map wrote (tests/golden/voc_map.npz, CPU part), and the on-device
matching kernel cvpp_voc_match against both (-m gpu part)."""
import os

import numpy as np
import pytest

import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ap_from_flags(g, flag, names):
    from computervision.pytorch_b200.core.metrics.mAP import truncated_confidence, voc_ap
    rows = g["det_rows"]
    n_gt = np.bincount(g["gt_cls"][g["gt_diff"] == 0], minlength=len(names))
    out = {}
    for c, name in enumerate(names):
        if n_gt[c] == 0:
            continue
        sel = np.nonzero(rows[:, 0].astype(int) == c)[0]
        order = np.argsort(-truncated_confidence(rows[sel, 1]), kind="stable")
        f = flag[sel][order]
        tp, fp = np.cumsum(f == 1), np.cumsum(f == 2)
        rec = [float(t) / max(int(n_gt[c]), 1) for t in tp]
        prec = [float(t) / max(int(p + t), 1) for t, p in zip(tp, fp)]
        out[name] = (voc_ap(rec, prec)[0], rec, prec)
    return out


@pytest.mark.parametrize("tag,thr", [("50", 0.5), ("75", 0.75)])
def test_oracle_matching_and_host_ap_reproduce_get_map(tag, thr):
    g = np.load(os.path.join(GOLD, "voc_map.npz"))
    names = [str(n) for n in g["names"]]
    flag = oracle.voc_match(g["det_rows"], g["det_counts"], g["gt_boxes"], g["gt_cls"], g["gt_diff"], g["gt_counts"], thr)
    res = _ap_from_flags(g, flag, names)
    classes = [str(c) for c in g[f"classes_{tag}"]]
    assert sorted(res) == classes and "bottle" not in res and "bus" not in res
    for name, ap in zip(classes, g[f"ap_{tag}"]):
        assert res[name][0] == ap                                    # Python-float arithmetic: bit for bit
        assert res[name][1] == list(g[f"rec_{tag}_{name}"]) and res[name][2] == list(g[f"prec_{tag}_{name}"])
    m = sum(r[0] for r in res.values()) / len(res)
    assert "metrics = {0:.2f}%".format(m * 100) == str(g[f"map_text_{tag}"])
    assert (flag == 0).sum() > 0 and (flag == 1).sum() > 20 and (flag == 2).sum() > 50


@pytest.mark.gpu
@pytest.mark.parametrize("tag,thr", [("50", 0.5), ("75", 0.75)])
def test_device_matching_equals_get_map(tag, thr):
    import torch
    from computervision.pytorch_b200 import ops
    from computervision.pytorch_b200.core.metrics import get_map_from_rows
    g = np.load(os.path.join(GOLD, "voc_map.npz"))
    names = [str(n) for n in g["names"]]
    res = get_map_from_rows(thr, g["det_rows"], g["det_counts"], g["gt_boxes"], g["gt_cls"], g["gt_diff"], g["gt_counts"],
                            names, device="cuda:0")
    classes = [str(c) for c in g[f"classes_{tag}"]]
    assert res["classes"] == classes
    for name, ap in zip(classes, g[f"ap_{tag}"]):
        assert res["ap"][name] == ap
        assert res["rec"][name] == list(g[f"rec_{tag}_{name}"]) and res["prec"][name] == list(g[f"prec_{tag}_{name}"])
    assert "metrics = {0:.2f}%".format(res["map"] * 100) == str(g[f"map_text_{tag}"])
    # flags and the overlaps themselves against the oracle's Python-float loop
    dev = "cuda:0"
    off = lambda c: torch.tensor(np.concatenate([[0], np.cumsum(c)]), dtype=torch.int32, device=dev)   # noqa: E731
    flag, best, ov = ops.voc_match(torch.from_numpy(g["det_rows"]).to(dev), off(g["det_counts"]),
                                   torch.from_numpy(g["gt_boxes"]).to(dev), torch.from_numpy(g["gt_cls"]).to(dev),
                                   torch.from_numpy(g["gt_diff"]).to(dev), off(g["gt_counts"]), thr)
    want = oracle.voc_match(g["det_rows"], g["det_counts"], g["gt_boxes"], g["gt_cls"], g["gt_diff"], g["gt_counts"], thr)
    assert np.array_equal(flag.cpu().numpy(), want)
    assert float(ov.max()) <= 1.0 and int((best >= 0).sum()) > 50


@pytest.mark.gpu
def test_device_matching_random_large_vs_oracle():
    """Random integer boxes, many repeated matches and exact ties of the overlap (first maximum must win)."""
    import torch
    from computervision.pytorch_b200 import ops
    rng = np.random.Generator(np.random.PCG64(11))
    B = 64
    det_counts = rng.integers(0, 400, B)
    gt_counts = rng.integers(0, 30, B)
    N, G = int(det_counts.sum()), int(gt_counts.sum())

    def boxes(n):
        x1, y1 = rng.integers(0, 60, n), rng.integers(0, 60, n)
        return np.stack([x1, y1, x1 + rng.integers(1, 40, n), y1 + rng.integers(1, 40, n)], 1).astype(np.float32)
    rows = np.zeros((N, 6), np.float32)
    rows[:, 0] = rng.integers(0, 4, N)
    rows[:, 1] = rng.uniform(0.001, 1, N)
    rows[:, 2:] = boxes(N)
    o = 0
    for n in det_counts:                      # (image, class) groups in descending score order
        blk = rows[o:o + n]
        rows[o:o + n] = blk[np.lexsort((-blk[:, 1], blk[:, 0]))]
        o += n
    gtb, gtc, gtd = boxes(G), rng.integers(0, 4, G).astype(np.int32), (rng.random(G) < 0.2).astype(np.int32)
    want = oracle.voc_match(rows, det_counts, gtb, gtc, gtd, gt_counts, 0.5)
    dev = "cuda:0"
    off = lambda c: torch.tensor(np.concatenate([[0], np.cumsum(c)]), dtype=torch.int32, device=dev)   # noqa: E731
    flag, _, _ = ops.voc_match(torch.from_numpy(rows).to(dev), off(det_counts), torch.from_numpy(gtb).to(dev),
                               torch.from_numpy(gtc).to(dev), torch.from_numpy(gtd).to(dev), off(gt_counts), 0.5)
    assert np.array_equal(flag.cpu().numpy(), want)
    assert (want == 1).sum() > 100 and (want == 0).sum() > 50
