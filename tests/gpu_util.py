"""Helpers shared by the -m gpu parity tests (they all go through the C ABI via ops.py)."""
import json
import os

import numpy as np
import torch

BOX_RTOL, BOX_ATOL, SCORE_RTOL = 1e-5, 2.5e-4, 1e-5   # see tests/test_oracle_golden.py


def record_error(got, ref, what=""):
    """CVPP_ERR_LOG=<file>: append the largest absolute and relative difference of this comparison (DESIGN.md quotes the
    measured maxima per path next to the tolerances the tests assert)."""
    path = os.environ.get("CVPP_ERR_LOG")
    if not path or np.size(ref) == 0:
        return
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    if got.shape != ref.shape:
        return
    d = np.abs(got - ref)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(np.abs(ref) > 0, d / np.abs(ref), 0.0)
    with open(path, "a") as f:
        f.write(json.dumps({"test": os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0], "what": what, "n": int(d.size),
                            "max_abs": float(d.max()), "max_rel": float(rel.max()), "max_ref": float(np.abs(ref).max())}) + "\n")


def to_dev(arrays, device="cuda:0"):
    return [torch.from_numpy(np.ascontiguousarray(a)).to(device) for a in arrays]


def decode_keys(key_row: np.ndarray):
    """uint64 class-major keys -> (cls, score, anchor)."""
    k = key_row.astype(np.uint64)
    cls = (k >> np.uint64(52)).astype(np.int64)
    inv = ((k >> np.uint64(21)) & np.uint64(0x7FFFFFFF)).astype(np.uint32)
    score = (np.uint32(0x7FFFFFFF) - inv).view(np.float32)
    anchor = (k & np.uint64(0x1FFFFF)).astype(np.int64)
    return cls, score, anchor


def candidates_to_numpy(c):
    """ops.Candidates -> per image (anchor-sorted) dict of arrays."""
    key = c.key.cpu().numpy().view(np.uint64)
    cnt = c.count.cpu().numpy()
    box = c.box_dense.cpu().numpy()
    out = []
    for b in range(key.shape[0]):
        n = int(cnt[b])
        assert n <= c.max_cand
        cls, score, anchor = decode_keys(key[b, :n])
        o = np.argsort(anchor, kind="stable")
        out.append(dict(cls=cls[o], score=score[o], anchor=anchor[o], box=box[b, anchor[o]]))
    return out


def assert_boxes_close(got, ref):
    assert got.shape == ref.shape
    record_error(got, ref, "box")
    assert np.all(np.abs(got - ref) <= BOX_RTOL * np.abs(ref) + BOX_ATOL), float(np.abs(got - ref).max())


def assert_scores_close(got, ref):
    assert got.shape == ref.shape
    record_error(got, ref, "score")
    assert np.all(np.abs(got - ref) <= SCORE_RTOL * np.abs(ref)), float((np.abs(got - ref) / np.abs(ref)).max())


def dets_to_numpy(det):
    cnt = det.count.cpu().numpy()
    box, score = det.box.cpu().numpy(), det.score.cpu().numpy()
    cls, anchor = det.cls.cpu().numpy(), det.anchor.cpu().numpy()
    out = []
    for b, n in enumerate(cnt):
        n = int(n)
        assert n <= box.shape[1]
        out.append(dict(box=box[b, :n], score=score[b, :n], cls=cls[b, :n], anchor=anchor[b, :n]))
    return out
