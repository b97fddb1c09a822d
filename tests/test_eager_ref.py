"""CPU: pins oracle/eager_gpu.py (the eager-torch restatement bench.py times on CUDA tensors as the "reference on
the GPU" bar) against the fixtures the REAL reference wrote: on CPU tensors it must reproduce them bit for bit -
same ops, same order, the same torchvision.ops.batched_nms call."""
import os

import numpy as np
import torch

import synth
from oracle import eager_gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _split(flat, counts):
    out, o = [], 0
    for c in counts:
        out.append(flat[o:o + c])
        o += c
    return out


def test_detect_tail_is_bit_identical_to_the_reference():
    g = np.load(os.path.join(GOLD, "yolov8_small.npz"))
    levels = [torch.from_numpy(g[f"level{i}"]) for i in range(3)]
    y = eager_gpu.detect_tail(levels, synth.YOLOV8_STRIDES, 80)
    assert np.array_equal(y.numpy(), g["y"])


def test_nms_rows_and_kept_anchors_are_bit_identical_to_the_reference():
    g = np.load(os.path.join(GOLD, "yolov8_small.npz"))
    y = torch.from_numpy(g["y"])
    for tag in "abc":
        conf, iou, md = g[f"params_{tag}"]
        rows, anchors = eager_gpu.non_max_suppression(y, float(conf), float(iou), int(md), nc=80)
        counts = g[f"counts_{tag}"]
        assert [len(a) for a in anchors] == list(counts)
        for r, a, rr, ra in zip(rows, anchors, _split(g[f"rows_{tag}"], counts), _split(g[f"anchors_{tag}"], counts)):
            assert np.array_equal(a.numpy(), ra) and np.array_equal(r.numpy(), rr)
