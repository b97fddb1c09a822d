"""CPU tests (-m "not gpu"): the host-side helpers of the mirror against fixtures produced by the
reference's own functions, and the C-ABI surface (library loads, exports every declared symbol)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from computervision.pytorch_b200 import _lib
from computervision.pytorch_b200.core.models.yolov8.modules import DFL
from computervision.pytorch_b200.core.utils import anchor, bboxes, image_process
from computervision.pytorch_b200.core.utils.ultralytics_ops import non_max_suppression, xywh2xyxy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "host_helpers.npz"))


def test_abi_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "cvpp.h")).read()
    declared = set(re.findall(r"CVPP_API\s+[\w\s\*]+?\b(cvpp_\w+)\s*\(", header))
    assert declared, "no declarations parsed from include/cvpp.h"
    lib = ctypes.CDLL(_lib.SO_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"libcvpp.so does not export {name}"
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())
    assert _lib.lib().cvpp_version() >= 100
    assert _lib.lib().cvpp_error_name(-3) == b"CVPP_ERR_WORKSPACE"
    assert _lib.lib().cvpp_yolov8_workspace_bytes(64, 8400, 8400, 80) > 0


def test_make_anchors(g):
    feats = [torch.zeros(1, 4, h, w) for h, w in ((6, 5), (3, 3), (2, 1))]
    ap, st = anchor.make_anchors(feats, [8, 16, 32], 0.5)
    assert np.array_equal(ap.numpy(), g["anchor_points"]) and np.array_equal(st.numpy(), g["anchor_strides"])


def test_box_converters(g):
    dist, pts = torch.from_numpy(g["d2b_dist"]), torch.from_numpy(g["d2b_pts"])
    assert np.array_equal(bboxes.dist2bbox(dist, pts, xywh=True, dim=1).numpy(), g["d2b_xywh"])
    assert np.array_equal(bboxes.dist2bbox(dist, pts, xywh=False, dim=1).numpy(), g["d2b_xyxy"])
    b = g["cv_in"]
    assert np.array_equal(xywh2xyxy(torch.from_numpy(b[:, :4])).numpy(), g["cv_xywh2xyxy"])
    assert np.array_equal(xywh2xyxy(b[:, :4]), g["cv_xywh2xyxy_np"])
    assert np.array_equal(bboxes.xywh_to_xyxy_torch(torch.from_numpy(b)).numpy(), g["cv_torch"])
    assert np.array_equal(bboxes.xywh_to_xyxy_torch(torch.from_numpy(b), more=True).numpy(), g["cv_torch_more"])
    assert np.array_equal(bboxes.xywh_to_xyxy(b[:, :4]), g["cv_numpy"])


def test_letterbox_reverse(g):
    b = g["cv_in"]
    for i, (h, w) in enumerate(g["lb_shapes"]):
        h, w = int(h), int(w)
        for inp in ((640, 640), (384, 512)):
            tag = f"{i}_{inp[0]}"
            f = image_process
            assert np.array_equal(f.reverse_letter_box_numpy([h, w], list(inp), b[:, :4].copy(), True), g[f"lb_np_xywh_{tag}"])
            assert np.array_equal(f.reverse_letter_box_numpy([h, w], list(inp), b[:, :4].copy(), False), g[f"lb_np_xyxy_{tag}"])
            assert np.array_equal(f.reverse_letter_box(h, w, list(inp), torch.from_numpy(b[:, :4].copy()), True).numpy(),
                                  g[f"lb_t_xywh_{tag}"])
            assert np.array_equal(f.reverse_letter_box(h, w, list(inp), torch.from_numpy(b[:, :4].copy()), False).numpy(),
                                  g[f"lb_t_xyxy_{tag}"])
            assert np.array_equal(f.yolo_correct_boxes(b[:, 0:2].copy(), b[:, 2:4].copy(), list(inp), [h, w], True),
                                  g[f"ycb_lb_{tag}"])
            assert np.array_equal(f.yolo_correct_boxes(b[:, 0:2].copy(), b[:, 2:4].copy(), list(inp), [h, w], False),
                                  g[f"ycb_nolb_{tag}"])


def test_anchor_tables(g):
    from types import SimpleNamespace as NS
    priors = anchor.generate_ssd_anchor_v2((300, 300), [30, 60, 111, 162, 213, 264, 315], [38, 19, 10, 5, 3, 1],
                                           [[1, 2, 1.0 / 2], [1, 2, 1.0 / 2, 3, 1.0 / 3], [1, 2, 1.0 / 2, 3, 1.0 / 3],
                                            [1, 2, 1.0 / 2, 3, 1.0 / 3], [1, 2, 1.0 / 2], [1, 2, 1.0 / 2]])
    ref = g["ssd_priors"]
    if priors.shape != ref.shape or not np.array_equal(priors, ref):
        pytest.fail(f"SSD priors differ: {priors.shape} vs {ref.shape}")
    c3 = NS(arch=NS(input_size=(3, 416, 416),
                    anchor=[116, 90, 156, 198, 373, 326, 30, 61, 62, 45, 59, 119, 10, 13, 16, 30, 33, 23]))
    assert np.array_equal(anchor.generate_yolo3_anchor(c3, None).numpy(), g["yolo3_anchors"])
    assert np.array_equal(anchor.generate_yolo3_anchor(c3, None, 1).numpy(), g["yolo3_anchors_1"])
    c7 = NS(arch=NS(anchors=[12, 16, 19, 36, 40, 28, 36, 75, 76, 55, 72, 146, 142, 110, 192, 243, 459, 401]))
    assert np.array_equal(anchor.get_yolo7_anchors(c7), g["yolo7_anchors"])


def test_dfl_module(g):
    out = DFL(16)(torch.from_numpy(g["dfl_in"])).numpy()
    assert np.allclose(out, g["dfl_out"], rtol=1e-5, atol=1e-5)


def test_nms_argument_errors_match_reference():
    p = torch.zeros((1, 84, 16))
    with pytest.raises(AssertionError, match="Invalid Confidence threshold"):
        non_max_suppression(p, conf_thres=1.5)
    with pytest.raises(AssertionError, match="Invalid IoU"):
        non_max_suppression(p, iou_thres=-0.1)
    with pytest.raises(ValueError, match="GPU only"):      # no CPU fallback
        non_max_suppression(p)
