"""Thin torch-facing wrappers over the C ABI: they own the device buffers (torch tensors) and pass raw
pointers + the current CUDA stream to libcvpp.  Nothing here computes on the CPU."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (ORDER_CLASS_MAJOR, ORDER_SCORE_DESC, RULE_COORD_TRICK, RULE_PER_CLASS,  # noqa: F401
                   RULE_TORCHVISION_CPU, check)

c_vp = ctypes.c_void_p


def _stream(device) -> c_vp:
    return c_vp(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> c_vp:
    return c_vp(0 if t is None else t.data_ptr())


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise ValueError(f"{what} must live on a CUDA device: libcvpp has no CPU path (got {t.device})")
    if t.dtype != torch.float32:
        raise ValueError(f"{what} must be float32 (got {t.dtype})")


@dataclass
class LevelSet:
    """Host-side description of the head levels as the C ABI wants it."""
    ptr: ctypes.Array
    batch_stride: ctypes.Array
    chan_stride: ctypes.Array
    h: ctypes.Array
    w: ctypes.Array
    stride: ctypes.Array
    n: int
    B: int
    C: int
    A: int
    device: torch.device
    keep: tuple  # tensors kept alive


def make_levels(levels: Sequence[torch.Tensor], strides: Sequence[float],
                sizes: Optional[Sequence[Tuple[int, int]]] = None) -> LevelSet:
    """levels: list of (B, C, H, W) tensors (any batch/channel strides, cells contiguous), or - with
    `sizes` - list of (B, C, H*W) views, e.g. slices of the concatenated x_cat."""
    n = len(levels)
    if n < 1 or n > 4 or len(strides) != n:
        raise ValueError("between 1 and 4 head levels with one stride each are supported")
    keep = []
    hs, ws = [], []
    for i, t in enumerate(levels):
        _require_cuda(t, f"level {i}")
        if sizes is None:
            if t.dim() != 4:
                raise ValueError(f"level {i}: expected (B, C, H, W), got {tuple(t.shape)}")
            H, W = int(t.shape[2]), int(t.shape[3])
            if t.stride(3) != 1 or t.stride(2) != W:
                t = t.contiguous()
        else:
            H, W = sizes[i]
            if t.dim() != 3 or t.shape[2] != H * W:
                raise ValueError(f"level {i}: expected (B, C, {H * W}), got {tuple(t.shape)}")
            if t.stride(2) != 1:
                t = t.contiguous()
        hs.append(H)
        ws.append(W)
        keep.append(t)
    B, C = int(keep[0].shape[0]), int(keep[0].shape[1])
    for t in keep:
        if t.shape[0] != B or t.shape[1] != C or t.device != keep[0].device:
            raise ValueError("all levels must share batch size, channel count and device")
    return LevelSet(
        ptr=(c_vp * n)(*[t.data_ptr() for t in keep]),
        batch_stride=(ctypes.c_int64 * n)(*[t.stride(0) for t in keep]),
        chan_stride=(ctypes.c_int64 * n)(*[t.stride(1) for t in keep]),
        h=(ctypes.c_int * n)(*hs), w=(ctypes.c_int * n)(*ws),
        stride=(ctypes.c_float * n)(*[float(s) for s in strides]),
        n=n, B=B, C=C, A=sum(h * w for h, w in zip(hs, ws)), device=keep[0].device, keep=tuple(keep))


@dataclass
class Candidates:
    key: torch.Tensor        # (B, max_cand) int64 (bit pattern of the uint64 key)
    count: torch.Tensor      # (B,) int32
    box_dense: torch.Tensor  # (B, A, 4) float32, valid at candidate anchors only
    max_cand: int
    A: int
    nc: int


@dataclass
class Detections:
    box: torch.Tensor     # (B, max_out, 4) xyxy
    score: torch.Tensor   # (B, max_out)
    cls: torch.Tensor     # (B, max_out) int32
    anchor: torch.Tensor  # (B, max_out) int32
    count: torch.Tensor   # (B,) int32
    cand_count: Optional[torch.Tensor] = None  # (B,) int32 candidates that entered NMS


def yolov8_decode_filter(ls: LevelSet, nc: int, conf_thres: float, reg_max: int = 16,
                         max_cand: Optional[int] = None) -> Candidates:
    if ls.C != 4 * reg_max + nc:
        raise ValueError(f"head has {ls.C} channels, expected 4*{reg_max}+{nc}")
    max_cand = int(max_cand or ls.A)
    dev = ls.device
    key = torch.empty((ls.B, max_cand), dtype=torch.int64, device=dev)
    count = torch.empty((ls.B,), dtype=torch.int32, device=dev)
    box_dense = torch.empty((ls.B, ls.A, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().cvpp_yolov8_decode_filter(ls.ptr, ls.batch_stride, ls.chan_stride, ls.h, ls.w, ls.stride, ls.n,
                                                   ls.B, nc, reg_max, float(conf_thres), _ptr(key), _ptr(count),
                                                   _ptr(box_dense), max_cand, _stream(dev)))
    return Candidates(key, count, box_dense, max_cand, ls.A, nc)


def yolov8_decode_full(ls: LevelSet, nc: int, reg_max: int = 16) -> torch.Tensor:
    if ls.C != 4 * reg_max + nc:
        raise ValueError(f"head has {ls.C} channels, expected 4*{reg_max}+{nc}")
    y = torch.empty((ls.B, 4 + nc, ls.A), dtype=torch.float32, device=ls.device)
    with torch.cuda.device(ls.device):
        check(_lib.lib().cvpp_yolov8_decode_full(ls.ptr, ls.batch_stride, ls.chan_stride, ls.h, ls.w, ls.stride, ls.n,
                                                 ls.B, nc, reg_max, _ptr(y), _stream(ls.device)))
    return y


def pred_filter(pred: torch.Tensor, nc: int, conf_thres: float, max_cand: Optional[int] = None) -> Candidates:
    _require_cuda(pred, "prediction")
    if pred.dim() != 3:
        raise ValueError(f"prediction must be (B, 4+nc+nm, A), got {tuple(pred.shape)}")
    pred = pred.contiguous()
    B, ch, A = (int(v) for v in pred.shape)
    max_cand = int(max_cand or A)
    dev = pred.device
    key = torch.empty((B, max_cand), dtype=torch.int64, device=dev)
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    box_dense = torch.empty((B, A, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().cvpp_pred_filter(_ptr(pred), B, ch, nc, A, float(conf_thres), _ptr(key), _ptr(count),
                                          _ptr(box_dense), max_cand, _stream(dev)))
    return Candidates(key, count, box_dense, max_cand, A, nc)


def segmented_sort(c: Candidates, rule: int = RULE_TORCHVISION_CPU, max_nms: int = 0) -> None:
    """Sorts c.key in place (and truncates c.count to max_nms when max_nms > 0)."""
    l = _lib.lib()
    B = int(c.key.shape[0])
    dev = c.key.device
    nbytes = int(l.cvpp_sort_workspace_bytes(B, c.max_cand))
    ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(l.cvpp_segmented_sort(_ptr(c.key), _ptr(c.count), B, c.max_cand, rule, int(max_nms), _ptr(ws), nbytes,
                                    _stream(dev)))


def nms(c: Candidates, iou_thres: float, rule: int = RULE_TORCHVISION_CPU, order: int = ORDER_SCORE_DESC,
        max_det: int = 300, max_out: Optional[int] = None) -> Detections:
    l = _lib.lib()
    B = int(c.key.shape[0])
    dev = c.key.device
    if max_out is None:
        max_out = max_det if (order == ORDER_SCORE_DESC and max_det > 0) else c.max_cand
    max_out = max(int(max_out), 1)
    nbytes = int(l.cvpp_nms_workspace_bytes(B, c.max_cand, c.nc))
    ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=dev)
    det = Detections(box=torch.empty((B, max_out, 4), dtype=torch.float32, device=dev),
                     score=torch.empty((B, max_out), dtype=torch.float32, device=dev),
                     cls=torch.empty((B, max_out), dtype=torch.int32, device=dev),
                     anchor=torch.empty((B, max_out), dtype=torch.int32, device=dev),
                     count=torch.empty((B,), dtype=torch.int32, device=dev), cand_count=c.count)
    with torch.cuda.device(dev):
        check(l.cvpp_nms(_ptr(c.key), _ptr(c.count), _ptr(c.box_dense), B, c.max_cand, c.A, c.nc, float(iou_thres),
                         rule, order, int(max_det), max_out, _ptr(det.box), _ptr(det.score), _ptr(det.cls),
                         _ptr(det.anchor), _ptr(det.count), _ptr(ws), nbytes, _stream(dev)))
    return det


class Yolov8Postprocessor:
    """Pre-allocated buffers + one C call (cvpp_yolov8_postprocess) per batch: decode+filter, sort, NMS.

    `capture()` records that call into a CUDA graph bound to one LevelSet (fixed input buffers), so a
    steady-state step is a single graph launch - the form to use for bs=1 latency."""

    def __init__(self, B: int, A: int, nc: int, device, max_det: int = 300, max_cand: Optional[int] = None):
        self.B, self.A, self.nc, self.max_det = int(B), int(A), int(nc), int(max_det)
        self.max_cand = int(max_cand or A)
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._lib = _lib.lib()
        self.ws_bytes = int(self._lib.cvpp_yolov8_workspace_bytes(self.B, self.A, self.max_cand, self.nc))
        dev = self.device
        self.ws = torch.empty((self.ws_bytes,), dtype=torch.uint8, device=dev)
        self.det = Detections(box=torch.empty((B, max_det, 4), dtype=torch.float32, device=dev),
                              score=torch.empty((B, max_det), dtype=torch.float32, device=dev),
                              cls=torch.empty((B, max_det), dtype=torch.int32, device=dev),
                              anchor=torch.empty((B, max_det), dtype=torch.int32, device=dev),
                              count=torch.empty((B,), dtype=torch.int32, device=dev),
                              cand_count=torch.empty((B,), dtype=torch.int32, device=dev))
        d = self.det
        self._out_args = (_ptr(d.box), _ptr(d.score), _ptr(d.cls), _ptr(d.anchor), _ptr(d.count), _ptr(d.cand_count),
                          _ptr(self.ws), self.ws_bytes)

    def __call__(self, ls: LevelSet, conf_thres: float, iou_thres: float, rule: int = RULE_TORCHVISION_CPU,
                 max_nms: int = 30000, reg_max: int = 16) -> Detections:
        if ls.B != self.B or ls.A != self.A or ls.C != 4 * reg_max + self.nc:
            raise ValueError("level set does not match the shapes this post-processor was built for")
        dev = self.device
        if torch.cuda.current_device() != dev.index:
            with torch.cuda.device(dev):
                return self.__call__(ls, conf_thres, iou_thres, rule, max_nms, reg_max)
        check(self._lib.cvpp_yolov8_postprocess(
            ls.ptr, ls.batch_stride, ls.chan_stride, ls.h, ls.w, ls.stride, ls.n, self.B, self.nc, reg_max,
            float(conf_thres), float(iou_thres), rule, self.max_det, int(max_nms), self.max_cand, *self._out_args,
            c_vp(torch.cuda.current_stream().cuda_stream)))
        return self.det

    def capture(self, ls: LevelSet, conf_thres: float, iou_thres: float, rule: int = RULE_TORCHVISION_CPU,
                max_nms: int = 30000, reg_max: int = 16) -> "GraphedPostprocess":
        return GraphedPostprocess(self, ls, (conf_thres, iou_thres, rule, max_nms, reg_max))


class GraphedPostprocess:
    """A CUDA-graph capture of Yolov8Postprocessor.__call__ on fixed input buffers."""

    def __init__(self, post: Yolov8Postprocessor, ls: LevelSet, args):
        self.post, self.ls = post, ls
        with torch.cuda.device(post.device):
            post(ls, *args)  # warm-up outside capture: one-time attribute / driver-entry-point work
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                post(ls, *args)

    def replay(self) -> Detections:
        self.graph.replay()
        return self.post.det


def letterbox_params(image_hw: Sequence[Tuple[int, int]], input_hw: Sequence[int], device) -> torch.Tensor:
    """(B, 5) fp32 rows in_w, in_h, left, top, scale: the scalars of reverse_letter_box
    (image_process.py:115-121) computed in Python doubles exactly like the reference, then cast."""
    rows = []
    for (h, w) in image_hw:
        scale = max(h / input_hw[0], w / input_hw[1])
        top = (input_hw[0] - h / scale) // 2
        left = (input_hw[1] - w / scale) // 2
        rows.append([float(input_hw[1]), float(input_hw[0]), left, top, scale])
    return torch.tensor(rows, dtype=torch.float64).to(torch.float32).to(device)


@dataclass
class CenterDetections:
    box: torch.Tensor     # (B, K, 4) xyxy (normalised, or original-image pixels with a letterbox)
    score: torch.Tensor   # (B, K)
    cls: torch.Tensor     # (B, K) int32
    pixel: torch.Tensor   # (B, K) int32  y*W + x of the peak
    count: torch.Tensor   # (B,) int32


def centernet_decode(pred: torch.Tensor, K: int, conf_thres: float, use_nms: bool = False, nms_thres: float = 0.5,
                     letterbox: Optional[torch.Tensor] = None, pool_mode: int = 0) -> CenterDetections:
    """pred (B, H, W, nc+4) NHWC -> per-image top-K peaks as boxes (cvpp_centernet_decode)."""
    _require_cuda(pred, "pred")
    if pred.dim() != 4 or pred.shape[3] < 5:
        raise ValueError(f"pred must be (B, H, W, nc+4), got {tuple(pred.shape)}")
    pred = pred.contiguous()
    B, H, W, Cf = (int(v) for v in pred.shape)
    nc = Cf - 4
    dev = pred.device
    l = _lib.lib()
    nbytes = int(l.cvpp_centernet_workspace_bytes(B, H, W, nc, K))
    ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=dev)
    out = CenterDetections(box=torch.empty((B, K, 4), dtype=torch.float32, device=dev),
                           score=torch.empty((B, K), dtype=torch.float32, device=dev),
                           cls=torch.empty((B, K), dtype=torch.int32, device=dev),
                           pixel=torch.empty((B, K), dtype=torch.int32, device=dev),
                           count=torch.empty((B,), dtype=torch.int32, device=dev))
    if letterbox is not None:
        letterbox = letterbox.to(device=dev, dtype=torch.float32).contiguous()
        if tuple(letterbox.shape) != (B, 5):
            raise ValueError("letterbox must be (B, 5)")
    with torch.cuda.device(dev):
        check(l.cvpp_centernet_decode(_ptr(pred), B, H, W, nc, int(K), float(conf_thres), int(pool_mode), int(use_nms),
                                      float(nms_thres), _ptr(letterbox), _ptr(out.box), _ptr(out.score), _ptr(out.cls),
                                      _ptr(out.pixel), _ptr(out.count), _ptr(ws), nbytes, _stream(dev)))
    return out


def diou_nms(boxes: torch.Tensor, scores: torch.Tensor, thr: float) -> torch.Tensor:
    """cvpp_diou_nms: int64 kept indices in descending score order."""
    _require_cuda(boxes, "boxes")
    _require_cuda(scores, "scores")
    boxes = boxes.reshape(-1, 4).contiguous()
    scores = scores.reshape(-1).contiguous()
    n = int(boxes.shape[0])
    dev = boxes.device
    keep = torch.empty((max(n, 1),), dtype=torch.int64, device=dev)
    cnt = torch.empty((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().cvpp_diou_nms(_ptr(boxes), _ptr(scores), n, float(thr), _ptr(keep), _ptr(cnt), _stream(dev)))
    return keep[: int(cnt.item())]


def ssd_decode_filter(loc: torch.Tensor, conf: torch.Tensor, priors: torch.Tensor, conf_thres: float,
                      max_cand: Optional[int] = None) -> Candidates:
    """loc (B,P,4), conf (B,P,nc+1) logits, priors (P,4) -> candidate keys per (prior, class-1)."""
    for t, name in ((loc, "loc"), (conf, "conf"), (priors, "priors")):
        _require_cuda(t, name)
    loc, conf, priors = loc.contiguous(), conf.contiguous(), priors.contiguous()
    B, P = int(loc.shape[0]), int(loc.shape[1])
    nc = int(conf.shape[2]) - 1
    if tuple(conf.shape[:2]) != (B, P) or tuple(priors.shape) != (P, 4) or loc.shape[2] != 4:
        raise ValueError("expected loc (B,P,4), conf (B,P,nc+1), priors (P,4)")
    max_cand = int(max_cand or P * nc)
    dev = loc.device
    key = torch.empty((B, max_cand), dtype=torch.int64, device=dev)
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    box_dense = torch.empty((B, P, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().cvpp_ssd_decode_filter(_ptr(loc), _ptr(conf), _ptr(priors), B, P, nc, float(conf_thres),
                                                _ptr(key), _ptr(count), _ptr(box_dense), max_cand, _stream(dev)))
    return Candidates(key, count, box_dense, max_cand, P, nc)


def ssd_parse_loc(loc: torch.Tensor, priors: torch.Tensor) -> torch.Tensor:
    """Ssd._parse_mbox_loc: loc (..., P, 4) + priors (P, 4) -> decoded, clamped xyxy of the same shape."""
    _require_cuda(loc, "loc")
    _require_cuda(priors, "priors")
    shape = loc.shape
    loc = loc.reshape(-1, shape[-2], 4).contiguous()
    priors = priors.contiguous()
    out = torch.empty_like(loc)
    with torch.cuda.device(loc.device):
        check(_lib.lib().cvpp_ssd_parse_loc(_ptr(loc), _ptr(priors), int(loc.shape[0]), int(loc.shape[1]), _ptr(out),
                                            _stream(loc.device)))
    return out.reshape(shape)


def per_class_nms_rows(c: Candidates, nms_thres: float, initial_out: int = 4096):
    """sort + per-class NMS (class-major order, no cap) and one device->host transfer.
    Returns per image (box (n,4), score (n,), cls (n,), anchor (n,)) as torch CPU tensors; retries with a
    larger output / candidate capacity when the first guess overflows."""
    B = int(c.key.shape[0])
    if int(c.count.max().item()) > c.max_cand:
        raise OverflowError("candidate buffer overflow: re-run the filter with a larger max_cand")
    segmented_sort(c, RULE_PER_CLASS)
    max_out = min(c.max_cand, initial_out)
    while True:
        det = nms(c, nms_thres, RULE_PER_CLASS, ORDER_CLASS_MAJOR, max_det=0, max_out=max_out)
        counts = det.count.cpu()
        if int(counts.max()) <= max_out:
            break
        max_out = int(counts.max())
    box, score, cls, anchor = det.box.cpu(), det.score.cpu(), det.cls.cpu(), det.anchor.cpu()
    return [(box[b, :n], score[b, :n], cls[b, :n], anchor[b, :n]) for b, n in enumerate(counts.tolist())]


def split_detections(det: Detections) -> List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]]:
    """One device->host read of the counts, then per-image views (box, score, cls, anchor)."""
    counts = det.count.tolist()
    cap = det.box.shape[1]
    out = []
    for b, n in enumerate(counts):
        if n > cap:
            raise RuntimeError(f"image {b}: {n} detections exceed the output capacity {cap}")
        out.append((det.box[b, :n], det.score[b, :n], det.cls[b, :n], det.anchor[b, :n]))
    return out
