"""CPU ORACLE — test infrastructure, not product code.

ctypes front-end of ``oracle/cvpp_oracle.c`` (the plain-C restatement of the
reference's decode / filter / NMS path) plus the few numpy restatements of
init-time tables.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package; ``computervision.pytorch_b200`` never does.

Parity status: the reference ships no tests or golden vectors, so the oracle is
pinned by fixtures generated from the *real* reference functions
(``tests/golden/make_golden.py``, run in the build container where
``/root/reference`` is mounted) and checked in ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcvpp_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force: bool = False) -> str:
    """Compile the C oracle with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "cvpp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_nms.restype = ctypes.c_int64
        _lib.orc_batched_nms.restype = ctypes.c_int64
        _lib.orc_nms_per_class.restype = ctypes.c_int64
        _lib.orc_max_threads.restype = ctypes.c_int
    return _lib


def set_threads(n: int) -> None:
    lib().orc_set_threads(ctypes.c_int(int(n)))


def max_threads() -> int:
    return int(lib().orc_max_threads())


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


# --------------------------------------------------------------------------
# torchvision restatements
# --------------------------------------------------------------------------
def nms(boxes, scores, iou_threshold: float) -> np.ndarray:
    """torchvision.ops.nms (CPU) restated; returns int64 kept indices, score-descending."""
    boxes = _f32(boxes).reshape(-1, 4)
    scores = _f32(scores).reshape(-1)
    n = boxes.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    k = lib().orc_nms(_ptr(boxes, _f32p), _ptr(scores, _f32p), ctypes.c_int64(n),
                      ctypes.c_double(float(iou_threshold)), _ptr(keep, _i64p))
    return keep[:k].copy()


def batched_nms(boxes, scores, idxs, iou_threshold: float, mode: int = 0) -> np.ndarray:
    """torchvision.ops.batched_nms (CPU) restated. mode 0=CPU rule, 1=trick, 2=vanilla, 3=the branch switch
    torchvision applies to CUDA tensors (numel > 20000 -> vanilla, boxes.py:80) with the CPU kernel's IoU test."""
    boxes = _f32(boxes).reshape(-1, 4)
    scores = _f32(scores).reshape(-1)
    idxs = _f32(idxs).reshape(-1)
    n = boxes.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    k = lib().orc_batched_nms(_ptr(boxes, _f32p), _ptr(scores, _f32p), _ptr(idxs, _f32p), ctypes.c_int64(n),
                              ctypes.c_double(float(iou_threshold)), ctypes.c_int(mode), _ptr(keep, _i64p))
    return keep[:k].copy()


def nms_per_class(boxes, scores, cls, nc: int, iou_threshold: float) -> np.ndarray:
    """Class-ascending loop of torchvision nms (YOLOv7._nms / Ssd.decode_boxes / yolo3_nms shape)."""
    boxes = _f32(boxes).reshape(-1, 4)
    scores = _f32(scores).reshape(-1)
    cls = np.ascontiguousarray(cls, dtype=np.int32).reshape(-1)
    n = boxes.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    k = lib().orc_nms_per_class(_ptr(boxes, _f32p), _ptr(scores, _f32p), _ptr(cls, _i32p), ctypes.c_int64(n),
                                ctypes.c_int(nc), ctypes.c_double(float(iou_threshold)), _ptr(keep, _i64p))
    return keep[:k].copy()


# --------------------------------------------------------------------------
# YOLOv8
# --------------------------------------------------------------------------
def yolov8_decode(levels: Sequence[np.ndarray], strides: Sequence[float], nc: int, reg_max: int = 16) -> np.ndarray:
    """Detect.forward eval tail (modules.py:434-445): list of (B,4*reg_max+nc,H,W) -> y (B,4+nc,A)."""
    levels = [_f32(l) for l in levels]
    B = levels[0].shape[0]
    nl = len(levels)
    hs = (ctypes.c_int * nl)(*[l.shape[2] for l in levels])
    ws = (ctypes.c_int * nl)(*[l.shape[3] for l in levels])
    st = (ctypes.c_float * nl)(*[float(s) for s in strides])
    ptrs = (_f32p * nl)(*[_ptr(l, _f32p) for l in levels])
    A = sum(l.shape[2] * l.shape[3] for l in levels)
    for l in levels:
        assert l.shape[1] == 4 * reg_max + nc, l.shape
    y = np.empty((B, 4 + nc, A), dtype=np.float32)
    lib().orc_yolov8_decode(ptrs, ctypes.c_int(nl), hs, ws, st, ctypes.c_int(B), ctypes.c_int(nc),
                            ctypes.c_int(reg_max), _ptr(y, _f32p))
    return y


def yolov8_nms(pred: np.ndarray, conf_thres: float, iou_thres: float, max_det: int = 300, nc: int = 0,
               max_nms: int = 30000, nms_mode: int = 0):
    """non_max_suppression (ultralytics_ops.py:131-264), live configuration.

    Returns (rows, anchors, cand_count): rows[b] is (n_b, 6) [x1,y1,x2,y2,conf,cls], anchors[b] the
    kept anchor indices (int32) in the same order.
    """
    pred = _f32(pred)
    B, ch, A = pred.shape
    nc = nc or (ch - 4)
    nm = ch - 4 - nc
    det = np.zeros((B, max_det, 6), dtype=np.float32)
    det_anchor = np.zeros((B, max_det), dtype=np.int32)
    det_count = np.zeros((B,), dtype=np.int32)
    cand_count = np.zeros((B,), dtype=np.int32)
    lib().orc_yolov8_nms(_ptr(pred, _f32p), ctypes.c_int(B), ctypes.c_int(nc), ctypes.c_int(nm), ctypes.c_int64(A),
                         ctypes.c_float(conf_thres), ctypes.c_double(float(iou_thres)), ctypes.c_int(max_det),
                         ctypes.c_int(max_nms), ctypes.c_int(nms_mode), _ptr(det, _f32p), _ptr(det_anchor, _i32p),
                         _ptr(det_count, _i32p), _ptr(cand_count, _i32p))
    rows = [det[b, :det_count[b]].copy() for b in range(B)]
    anchors = [det_anchor[b, :det_count[b]].copy() for b in range(B)]
    return rows, anchors, cand_count


def yolov8_candidates(pred: np.ndarray, conf_thres: float, nc: int = 0):
    """Candidates entering NMS, in anchor order: per image (boxes xyxy, score, cls, anchor)."""
    pred = _f32(pred)
    B, ch, A = pred.shape
    nc = nc or (ch - 4)
    nm = ch - 4 - nc
    box = np.zeros((B, A, 4), dtype=np.float32)
    score = np.zeros((B, A), dtype=np.float32)
    cls = np.zeros((B, A), dtype=np.int32)
    anc = np.zeros((B, A), dtype=np.int32)
    cnt = np.zeros((B,), dtype=np.int32)
    lib().orc_yolov8_candidates(_ptr(pred, _f32p), ctypes.c_int(B), ctypes.c_int(nc), ctypes.c_int(nm),
                                ctypes.c_int64(A), ctypes.c_float(conf_thres), _ptr(box, _f32p), _ptr(score, _f32p),
                                _ptr(cls, _i32p), _ptr(anc, _i32p), _ptr(cnt, _i32p))
    return [(box[b, :cnt[b]].copy(), score[b, :cnt[b]].copy(), cls[b, :cnt[b]].copy(), anc[b, :cnt[b]].copy())
            for b in range(B)]


# --------------------------------------------------------------------------
# CenterNet
# --------------------------------------------------------------------------
def letterbox_params(image_hw, input_hw) -> np.ndarray:
    """(in_w, in_h, left, top, scale) as float32, computed in Python doubles like
    reverse_letter_box (image_process.py:115-121)."""
    out = []
    for (h, w) in image_hw:
        scale = max(h / input_hw[0], w / input_hw[1])
        top = (input_hw[0] - h / scale) // 2
        left = (input_hw[1] - w / scale) // 2
        out.append([input_hw[1], input_hw[0], left, top, scale])
    return np.asarray(out, dtype=np.float32)


def diou_nms(boxes, scores, thr: float) -> np.ndarray:
    """core/utils/nms.py:9-31 restated; kept indices in descending score order."""
    boxes = _f32(boxes).reshape(-1, 4)
    scores = _f32(scores).reshape(-1)
    n = boxes.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    lib().orc_diou_nms.restype = ctypes.c_int64
    k = lib().orc_diou_nms(_ptr(boxes, _f32p), _ptr(scores, _f32p), ctypes.c_int64(n), ctypes.c_float(thr),
                           _ptr(keep, _i64p))
    return keep[:k].copy()


def centernet_decode(pred, K: int, conf: float, pool_mode: int = 0, use_nms: bool = False, nms_thr: float = 0.5,
                     letterbox=None):
    """CenterNetA.decode_boxes (centernet.py:271-314) per image. pred (B,H,W,nc+4) NHWC.
    Returns per image (boxes (n,4), scores (n,), classes (n,) int32, pixel (n,) int32)."""
    pred = _f32(pred)
    B, H, W, Cf = pred.shape
    nc = Cf - 4
    box = np.zeros((B, K, 4), np.float32)
    score = np.zeros((B, K), np.float32)
    cls = np.zeros((B, K), np.int32)
    pix = np.zeros((B, K), np.int32)
    cnt = np.zeros((B,), np.int32)
    lb = None if letterbox is None else _f32(letterbox)
    lib().orc_centernet_decode(_ptr(pred, _f32p), ctypes.c_int(B), ctypes.c_int(H), ctypes.c_int(W), ctypes.c_int(nc),
                               ctypes.c_int(K), ctypes.c_float(conf), ctypes.c_int(pool_mode), ctypes.c_int(int(use_nms)),
                               ctypes.c_float(nms_thr), None if lb is None else _ptr(lb, _f32p), _ptr(box, _f32p),
                               _ptr(score, _f32p), _ptr(cls, _i32p), _ptr(pix, _i32p), _ptr(cnt, _i32p))
    return [(box[b, :cnt[b]].copy(), score[b, :cnt[b]].copy(), cls[b, :cnt[b]].copy(), pix[b, :cnt[b]].copy())
            for b in range(B)]


def topk(scores, k: int):
    """torch.topk(scores.view(B, -1), k, largest=True, sorted=True) (centernet.py:330) with the product's tie rule
    made explicit: equal scores are ordered by the lower flat index (stable descending sort).
    Returns (values (B, k) float32, indices (B, k) int64)."""
    s = _f32(scores)
    s = s.reshape(s.shape[0], -1)
    idx = np.argsort(-s.astype(np.float64), axis=1, kind="stable")[:, :k].astype(np.int64)
    return np.take_along_axis(s, idx, axis=1), idx


def voc_match(det_rows, det_counts, gt_boxes, gt_cls, gt_diff, gt_counts, min_overlap: float):
    """The matching loop of get_map (core/metrics/mAP.py:441-520) restated per image in Python floats:
    det_rows (N, 6) VOC rows in file / line order -> flag (N,) 1 = TP, 2 = FP, 0 = matched a difficult box.
    Within an image the processing order of a class is the line order (module docstring of csrc/map_match.cu)."""
    det_rows = np.asarray(det_rows, np.float32).reshape(-1, 6)
    flag = np.zeros(len(det_rows), np.int32)
    d0 = g0 = 0
    for nd, ng in zip(det_counts, gt_counts):
        used = [False] * int(ng)
        for d in range(d0, d0 + int(nd)):
            c = int(det_rows[d, 0])
            bb = [float(v) for v in det_rows[d, 2:6]]
            ovmax, match = -1, -1
            for g in range(int(ng)):
                if int(gt_cls[g0 + g]) != c:
                    continue
                bbgt = [float(v) for v in gt_boxes[g0 + g]]
                bi = [max(bb[0], bbgt[0]), max(bb[1], bbgt[1]), min(bb[2], bbgt[2]), min(bb[3], bbgt[3])]
                iw = bi[2] - bi[0] + 1
                ih = bi[3] - bi[1] + 1
                if iw > 0 and ih > 0:
                    ua = (bb[2] - bb[0] + 1) * (bb[3] - bb[1] + 1) + (bbgt[2] - bbgt[0] + 1) * (bbgt[3] - bbgt[1] + 1) - iw * ih
                    ov = iw * ih / ua
                    if ov > ovmax:
                        ovmax, match = ov, g
            if ovmax >= min_overlap:
                if not int(gt_diff[g0 + match]):
                    if not used[match]:
                        flag[d] = 1
                        used[match] = True
                    else:
                        flag[d] = 2
            else:
                flag[d] = 2
        d0 += int(nd)
        g0 += int(ng)
    return flag


# --------------------------------------------------------------------------
# SSD
# --------------------------------------------------------------------------
def ssd_priors(input_hw=(300, 300), anchor_sizes=(30, 60, 111, 162, 213, 264, 315), feature_shapes=(38, 19, 10, 5, 3, 1),
               aspect_ratios=((1, 2, 0.5), (1, 2, 0.5, 3, 1.0 / 3), (1, 2, 0.5, 3, 1.0 / 3), (1, 2, 0.5, 3, 1.0 / 3),
                              (1, 2, 0.5), (1, 2, 0.5))) -> np.ndarray:
    """Ssd._get_ssd_anchors (ssd.py:482-541) restated: float64 numpy, cast to float32 at the end."""
    image_h, image_w = input_hw
    out = []
    for i, fs in enumerate(feature_shapes):
        lo, hi = anchor_sizes[i], anchor_sizes[i + 1]
        bw, bh = [], []
        for ar in aspect_ratios[i]:
            if ar == 1:
                bw += [lo, np.sqrt(lo * hi)]
                bh += [lo, np.sqrt(lo * hi)]
            else:
                bw.append(lo * np.sqrt(ar))
                bh.append(lo / np.sqrt(ar))
        hw_, hh_ = np.array(bw) / 2.0, np.array(bh) / 2.0
        px = [image_h / fs, image_w / fs]
        cx = np.linspace(0.5 * px[1], image_w - 0.5 * px[1], fs)
        cy = np.linspace(0.5 * px[0], image_h - 0.5 * px[0], fs)
        gx, gy = np.meshgrid(cx, cy)
        a = np.concatenate((gx.reshape(-1, 1), gy.reshape(-1, 1)), axis=1)
        a = np.tile(a, (1, len(bw) * 2))
        a[:, ::4] -= hw_
        a[:, 1::4] -= hh_
        a[:, 2::4] += hw_
        a[:, 3::4] += hh_
        a[:, ::2] /= image_w
        a[:, 1::2] /= image_h
        out.append(np.clip(a, 0.0, 1.0).reshape(-1, 4))
    return np.concatenate(out, axis=0).astype(np.float32)


def ssd_decode(loc, conf, priors, conf_thres: float, nms_thres: float = 0.5, cap: int = 0):
    """Ssd.decode_boxes (ssd.py:236-288) before the letterbox epilogue: per image rows (n, 6)
    [x1,y1,x2,y2,label,conf] in normalised coordinates, class-ascending then score-descending, and the
    prior index of every row."""
    loc, conf, priors = _f32(loc), _f32(conf), _f32(priors)
    B, P, _ = loc.shape
    nc = conf.shape[2] - 1
    cap = cap or P * nc
    rows = np.zeros((B, cap, 6), np.float32)
    prior = np.zeros((B, cap), np.int32)
    cnt = np.zeros((B,), np.int32)
    lib().orc_ssd_decode(_ptr(loc, _f32p), _ptr(conf, _f32p), _ptr(priors, _f32p), ctypes.c_int(B), ctypes.c_int64(P),
                         ctypes.c_int(nc), ctypes.c_float(conf_thres), ctypes.c_double(float(nms_thres)),
                         ctypes.c_int(cap), _ptr(rows, _f32p), _ptr(prior, _i32p), _ptr(cnt, _i32p))
    return [(rows[b, :min(cnt[b], cap)].copy(), prior[b, :min(cnt[b], cap)].copy()) for b in range(B)]


def yolo_correct_rows(rows: np.ndarray, input_hw, image_hw, letterbox: bool = True) -> np.ndarray:
    """The shared epilogue of Ssd.decode_boxes (:282-287) / YOLOv7._nms (:416-421): xyxy -> (centre, size) ->
    yolo_correct_boxes (image_process.py:161-181), float32 numpy with Python-double scalars."""
    out = rows.copy()
    if out.shape[0] == 0:
        return out
    xy = (out[:, 0:2] + out[:, 2:4]) / 2
    wh = out[:, 2:4] - out[:, 0:2]
    xywh = np.concatenate([xy, wh], axis=-1)
    if letterbox:
        nb = np.concatenate((xywh[..., 0:2] - xywh[..., 2:4] / 2, xywh[..., 0:2] + xywh[..., 2:4] / 2), axis=-1)
        nb[..., ::2] *= input_hw[1]
        nb[..., 1::2] *= input_hw[0]
        scale = max(image_hw[0] / input_hw[0], image_hw[1] / input_hw[1])
        top = (input_hw[0] - image_hw[0] / scale) // 2
        left = (input_hw[1] - image_hw[1] / scale) // 2
        nb[..., 0] -= left
        nb[..., 2] -= left
        nb[..., 1] -= top
        nb[..., 3] -= top
        nb *= scale
    else:
        nb = np.concatenate((xywh[..., 0:2] - xywh[..., 2:4] / 2, xywh[..., 0:2] + xywh[..., 2:4] / 2), axis=-1)
        nb[:, ::2] *= image_hw[1]
        nb[:, 1::2] *= image_hw[0]
    out[:, :4] = nb
    return out


# --------------------------------------------------------------------------
# YOLOv7
# --------------------------------------------------------------------------
YOLOV7_ANCHORS = (12, 16, 19, 36, 40, 28, 36, 75, 76, 55, 72, 146, 142, 110, 192, 243, 459, 401)
YOLOV7_MASK = ((6, 7, 8), (3, 4, 5), (0, 1, 2))


def yolov7_level_anchors(anchors=YOLOV7_ANCHORS, mask=YOLOV7_MASK) -> np.ndarray:
    """(num_levels*3, 2) anchors in pixels, row 3*i + a = anchors[mask[i][a]] (yolo_v7.py:259-262)."""
    a = np.array(anchors, dtype=np.float32).reshape(-1, 2)
    return np.concatenate([a[list(m)] for m in mask], axis=0)


def yolov7_decode(levels, nc: int, input_hw=(640, 640), anchors=YOLOV7_ANCHORS, mask=YOLOV7_MASK) -> np.ndarray:
    """YOLOv7.decode_box up to `decoded_outputs` (yolo_v7.py:245-344): (B, A, 5+nc)."""
    levels = [_f32(l) for l in levels]
    nl = len(levels)
    B = levels[0].shape[0]
    hs = (ctypes.c_int * nl)(*[l.shape[2] for l in levels])
    ws = (ctypes.c_int * nl)(*[l.shape[3] for l in levels])
    ptrs = (_f32p * nl)(*[_ptr(l, _f32p) for l in levels])
    la = yolov7_level_anchors(anchors, mask)
    A = sum(3 * l.shape[2] * l.shape[3] for l in levels)
    out = np.empty((B, A, 5 + nc), np.float32)
    lib().orc_yolov7_decode(ptrs, ctypes.c_int(nl), hs, ws, _ptr(la, _f32p), ctypes.c_int(input_hw[0]),
                            ctypes.c_int(input_hw[1]), ctypes.c_int(B), ctypes.c_int(nc), _ptr(out, _f32p))
    return out


def yolov7_nms(decoded, conf_thres: float, nms_thres: float, cap: int = 0):
    """YOLOv7._nms (yolo_v7.py:348-413) before the letterbox epilogue: per image (rows (n,7), anchors (n,))
    plus the candidate counts."""
    decoded = _f32(decoded)
    B, A, attrs = decoded.shape
    nc = attrs - 5
    cap = cap or A
    rows = np.zeros((B, cap, 7), np.float32)
    anc = np.zeros((B, cap), np.int32)
    cnt = np.zeros((B,), np.int32)
    cand = np.zeros((B,), np.int32)
    lib().orc_yolov7_nms(_ptr(decoded, _f32p), ctypes.c_int(B), ctypes.c_int64(A), ctypes.c_int(nc),
                         ctypes.c_float(conf_thres), ctypes.c_double(float(nms_thres)), ctypes.c_int(cap),
                         _ptr(rows, _f32p), _ptr(anc, _i32p), _ptr(cnt, _i32p), _ptr(cand, _i32p))
    return [(rows[b, :min(cnt[b], cap)].copy(), anc[b, :min(cnt[b], cap)].copy()) for b in range(B)], cand


# --------------------------------------------------------------------------
# YOLOv3
# --------------------------------------------------------------------------
YOLOV3_ANCHORS = (116, 90, 156, 198, 373, 326, 30, 61, 62, 45, 59, 119, 10, 13, 16, 30, 33, 23)


def yolov3_anchors_norm(anchors=YOLOV3_ANCHORS, input_hw=(416, 416)) -> np.ndarray:
    """generate_yolo3_anchor (core/utils/anchor.py:102-117): fp32 anchors divided in place by w, h."""
    a = np.array(anchors, dtype=np.float32).reshape(-1, 2)
    a[:, 0] /= np.float32(input_hw[1])
    a[:, 1] /= np.float32(input_hw[0])
    return a


def yolov3_dense(levels, nc: int, anchors=YOLOV3_ANCHORS, input_hw=(416, 416)):
    """Decoder.__call__ up to the NMS (yolov3_decode.py:40-63): boxes (M, 4), scores (M, nc) with the batch
    flattened per scale and the scales concatenated."""
    an = yolov3_anchors_norm(anchors, input_hw)
    boxes, scores = [], []
    for i, l in enumerate(levels):
        l = _f32(l)
        N, _, H, W = l.shape
        b = np.empty((N * H * W * 3, 4), np.float32)
        s = np.empty((N * H * W * 3, nc), np.float32)
        a = np.ascontiguousarray(an[3 * i:3 * i + 3])
        lib().orc_yolov3_scale(_ptr(l, _f32p), ctypes.c_int(N), ctypes.c_int(nc), ctypes.c_int(H), ctypes.c_int(W),
                               _ptr(a, _f32p), _ptr(b, _f32p), _ptr(s, _f32p))
        boxes.append(b)
        scores.append(s)
    return np.concatenate(boxes, 0), np.concatenate(scores, 0)


def yolo3_nms(boxes, scores, conf_thres: float, iou_thres: float):
    """yolo3_nms (core/utils/nms.py:54-84): -> (boxes (K,4), scores (K,), classes (K,) int32, rows (K,) int64,
    candidate count)."""
    boxes, scores = _f32(boxes), _f32(scores)
    M, nc = scores.shape
    l = lib()
    l.orc_yolo3_nms.restype = ctypes.c_int64
    cap = max(M, 1)
    while True:
        ob = np.zeros((cap, 4), np.float32)
        os_ = np.zeros((cap,), np.float32)
        oc = np.zeros((cap,), np.int32)
        orow = np.zeros((cap,), np.int64)
        cands = ctypes.c_int64(0)
        k = int(l.orc_yolo3_nms(_ptr(boxes, _f32p), _ptr(scores, _f32p), ctypes.c_int64(M), ctypes.c_int(nc),
                                ctypes.c_float(conf_thres), ctypes.c_double(float(iou_thres)), ctypes.c_int64(cap),
                                _ptr(ob, _f32p), _ptr(os_, _f32p), _ptr(oc, _i32p), _ptr(orow, _i64p),
                                ctypes.byref(cands)))
        if k <= cap:
            return ob[:k], os_[:k], oc[:k], orow[:k], int(cands.value)
        cap = k
