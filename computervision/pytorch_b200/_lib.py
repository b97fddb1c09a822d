"""ctypes binding of libcvpp.so (include/cvpp.h).  No CPU fallback: if the library is missing the
import of any op fails loudly."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libcvpp.so")

CVPP_OK = 0
RULE_TORCHVISION_CPU, RULE_COORD_TRICK, RULE_PER_CLASS, RULE_TORCHVISION_CUDA = 0, 1, 2, 3
ORDER_SCORE_DESC, ORDER_CLASS_MAJOR = 0, 1
ROWS_YOLOV8, ROWS_SSD, ROWS_YOLOV7, ROWS_FULL, ROWS_COCO, ROWS_VOC = 0, 1, 2, 3, 4, 5
BOX_KEEP, BOX_CORRECT, BOX_NORMALISE_CORRECT = 0, 1, 2

c_int, c_i64, c_f32, c_f64 = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double
c_vp, c_size = ctypes.c_void_p, ctypes.c_size_t
P = ctypes.POINTER


class CvppError(RuntimeError):
    """A libcvpp call returned a CVPP_ERR_* code."""

    def __init__(self, code: int, name: str, msg: str):
        super().__init__(f"{name} ({code}): {msg}")
        self.code = code
        self.name = name
        self.msg = msg


_PROTOS = {
    "cvpp_version": (c_int, []),
    "cvpp_last_error": (ctypes.c_char_p, []),
    "cvpp_error_name": (ctypes.c_char_p, [c_int]),
    "cvpp_yolov8_decode_filter": (c_int, [P(c_vp), P(c_i64), P(c_i64), P(c_int), P(c_int), P(c_f32), c_int, c_int,
                                          c_int, c_int, c_f32, c_vp, c_vp, c_vp, c_int, c_vp]),
    "cvpp_yolov8_head_decode_filter": (c_int, [P(c_vp), P(c_vp), P(c_vp), P(c_vp), P(c_vp), P(c_vp), P(c_int), P(c_int),
                                               P(c_f32), c_int, c_int, c_int, c_int, c_int, c_int, c_f32, c_vp, c_vp, c_vp,
                                               c_int, c_vp]),
    "cvpp_yolov8_head_decode_filter_x": (c_int, [P(c_vp), P(c_vp), P(c_vp), P(c_vp), P(c_vp), P(c_vp), P(c_int), P(c_int),
                                                 P(c_f32), c_int, c_int, c_int, c_int, c_int, c_int, c_f32, c_vp, c_vp, c_vp,
                                                 c_int, c_vp, c_vp]),
    "cvpp_yolov8_decode_full": (c_int, [P(c_vp), P(c_i64), P(c_i64), P(c_int), P(c_int), P(c_f32), c_int, c_int,
                                        c_int, c_int, c_vp, c_vp]),
    "cvpp_pred_filter": (c_int, [c_vp, c_int, c_int, c_int, c_i64, c_f32, c_vp, c_vp, c_vp, c_int, c_vp]),
    "cvpp_keep_classes": (c_int, [c_vp, c_vp, c_int, c_int, P(ctypes.c_int32), c_int, c_vp]),
    "cvpp_sort_workspace_bytes": (c_size, [c_int, c_int]),
    "cvpp_segmented_sort": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_size, c_vp]),
    "cvpp_nms_workspace_bytes": (c_size, [c_int, c_int, c_int]),
    "cvpp_nms": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_int, c_f64, c_int, c_int, c_int, c_int, c_vp, c_vp,
                         c_vp, c_vp, c_vp, c_vp, c_size, c_vp]),
    "cvpp_sort_nms_workspace_bytes": (c_size, [c_int, c_int, c_int]),
    "cvpp_sort_nms": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_int, c_f64, c_int, c_int, c_int, c_int, c_int,
                              c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_size, c_vp]),
    "cvpp_yolov8_workspace_bytes": (c_size, [c_int, c_i64, c_int, c_int]),
    "cvpp_yolov8_postprocess": (c_int, [P(c_vp), P(c_i64), P(c_i64), P(c_int), P(c_int), P(c_f32), c_int, c_int,
                                        c_int, c_int, c_f32, c_f64, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp,
                                        c_vp, c_vp, c_vp, c_vp, c_size, c_vp]),
    "cvpp_yolov8_postprocess_gather": (c_int, [P(c_vp), P(c_i64), P(c_i64), P(c_int), P(c_int), P(c_f32), c_int, c_int,
                                               c_int, c_int, c_f32, c_f64, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp,
                                               c_vp, c_vp, c_vp, c_vp, c_size, c_vp, P(c_vp), c_vp, c_int, c_int, c_vp]),
    "cvpp_yolov8_postprocess_ev": (c_int, [P(c_vp), P(c_i64), P(c_i64), P(c_int), P(c_int), P(c_f32), c_int, c_int,
                                           c_int, c_int, c_f32, c_f64, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp,
                                           c_vp, c_vp, c_vp, c_vp, c_size, c_vp, c_vp]),
    "cvpp_centernet_workspace_bytes": (c_size, [c_int, c_int, c_int, c_int, c_int]),
    "cvpp_centernet_decode": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_f32, c_int, c_int, c_f32, c_vp, c_vp,
                                      c_vp, c_vp, c_vp, c_vp, c_vp, c_size, c_vp]),
    "cvpp_diou_nms": (c_int, [c_vp, c_vp, c_int, c_f32, c_vp, c_vp, c_vp]),
    "cvpp_ssd_decode_filter": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_f32, c_vp, c_vp, c_vp, c_int, c_vp]),
    "cvpp_ssd_parse_loc": (c_int, [c_vp, c_vp, c_int, c_int, c_vp, c_vp]),
    "cvpp_yolov7_decode_filter": (c_int, [P(c_vp), P(c_i64), P(c_i64), P(c_int), P(c_int), P(c_f32), c_int, c_int,
                                          c_int, c_int, c_int, c_f32, c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    "cvpp_yolov7_pred_filter": (c_int, [c_vp, c_int, c_i64, c_int, c_f32, c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    "cvpp_yolov3_decode_filter": (c_int, [P(c_vp), P(c_i64), P(c_i64), P(c_int), P(c_int), P(c_f32), c_int, c_int,
                                          c_int, c_int, c_int, c_f32, c_int, c_vp, c_vp, c_vp, c_int, c_vp]),
    "cvpp_yolov3_predict_bbox": (c_int, [c_vp, c_int, c_int, c_int, c_int, P(c_f32), c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cvpp_score_matrix_filter": (c_int, [c_vp, c_i64, c_int, c_f32, c_vp, c_vp, c_int, c_vp]),
    "cvpp_gather_feat": (c_int, [c_vp, c_vp, c_int, c_vp, c_int, c_i64, c_int, c_int, c_vp, c_vp, c_vp]),
    "cvpp_letterbox_reverse": (c_int, [c_vp, c_i64, c_int, c_f32, c_f32, c_f32, c_f32, c_f32, c_vp, c_vp]),
    "cvpp_centernet_suppress": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "cvpp_detection_epilogue_allgather": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_int, c_int,
                                                  c_vp, P(c_vp), c_int, c_int, c_vp]),
    "cvpp_detection_epilogue_multicast": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_int, c_int,
                                                  c_vp, c_vp, c_int, c_int, c_vp]),
    "cvpp_detection_epilogue_compact": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_int, c_int,
                                                c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "cvpp_voc_match": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_f64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cvpp_topk_workspace_bytes": (c_size, [c_int, c_int]),
    "cvpp_topk": (c_int, [c_vp, c_int, c_i64, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_size, c_vp]),
    "cvpp_detection_epilogue": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_int, c_int, c_vp,
                                        c_vp, c_vp, c_vp]),
}

_lib = None


def exported_symbols():
    """Names declared in include/cvpp.h that this binding expects (used by the CPU-side ABI test)."""
    return sorted(_PROTOS)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C computervision/pytorch_b200/csrc`). There is no CPU fallback.")
        l = ctypes.CDLL(SO_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != CVPP_OK:
        l = lib()
        raise CvppError(rc, l.cvpp_error_name(rc).decode(), l.cvpp_last_error().decode())
