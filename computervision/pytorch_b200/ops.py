"""Thin torch-facing wrappers over the C ABI: they own the device buffers (torch tensors) and pass raw
pointers + the current CUDA stream to libcvpp.  Nothing here computes on the CPU."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (BOX_CORRECT, BOX_KEEP, BOX_NORMALISE_CORRECT, ORDER_CLASS_MAJOR, ORDER_SCORE_DESC,  # noqa: F401
                   ROWS_COCO, ROWS_FULL, ROWS_SSD, ROWS_VOC, ROWS_YOLOV7, ROWS_YOLOV8, RULE_COORD_TRICK, RULE_PER_CLASS,
                   RULE_TORCHVISION_CPU, RULE_TORCHVISION_CUDA, check)

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


def make_levels(levels: Sequence[torch.Tensor], strides: Optional[Sequence[float]] = None,
                sizes: Optional[Sequence[Tuple[int, int]]] = None) -> LevelSet:
    """levels: list of (B, C, H, W) tensors (any batch/channel strides, cells contiguous), or - with
    `sizes` - list of (B, C, H*W) views, e.g. slices of the concatenated x_cat.  `strides` is the YOLOv8
    stride per level (None for the anchor-based heads, which take an anchor table instead)."""
    n = len(levels)
    if strides is None:
        strides = [0.0] * n
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
    aux_dense: Optional[torch.Tensor] = None  # YOLOv7: (B, A, 2) float32 (obj, class_conf)


@dataclass
class Detections:
    box: torch.Tensor     # (B, max_out, 4) xyxy
    score: torch.Tensor   # (B, max_out)
    cls: torch.Tensor     # (B, max_out) int32
    anchor: torch.Tensor  # (B, max_out) int32
    count: torch.Tensor   # (B,) int32
    cand_count: Optional[torch.Tensor] = None  # (B,) int32 candidates that entered NMS


def _check_one_key_per_anchor(max_cand: int, A: int) -> None:
    """Filters that emit at most one key per anchor: a smaller candidate buffer could overflow, and keys past max_cand are
    dropped in the (nondeterministic) order of the atomics - wrong detections without an error."""
    if max_cand < A:
        raise ValueError(f"max_cand={max_cand} < A={A}: every anchor can pass the confidence filter; an overflowing candidate "
                         "buffer would silently drop keys in a nondeterministic order")


def yolov8_decode_filter(ls: LevelSet, nc: int, conf_thres: float, reg_max: int = 16,
                         max_cand: Optional[int] = None) -> Candidates:
    if ls.C != 4 * reg_max + nc:
        raise ValueError(f"head has {ls.C} channels, expected 4*{reg_max}+{nc}")
    max_cand = int(max_cand or ls.A)
    _check_one_key_per_anchor(max_cand, ls.A)
    dev = ls.device
    key = torch.empty((ls.B, max_cand), dtype=torch.int64, device=dev)
    count = torch.empty((ls.B,), dtype=torch.int32, device=dev)
    box_dense = torch.empty((ls.B, ls.A, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().cvpp_yolov8_decode_filter(ls.ptr, ls.batch_stride, ls.chan_stride, ls.h, ls.w, ls.stride, ls.n,
                                                   ls.B, nc, reg_max, float(conf_thres), _ptr(key), _ptr(count),
                                                   _ptr(box_dense), max_cand, _stream(dev)))
    return Candidates(key, count, box_dense, max_cand, ls.A, nc)


def yolov8_head_decode_filter(box_feats: Sequence[torch.Tensor], cls_feats: Sequence[torch.Tensor],
                              box_w: Sequence[torch.Tensor], box_b: Sequence[torch.Tensor], cls_w: Sequence[torch.Tensor],
                              cls_b: Sequence[torch.Tensor], strides: Sequence[float], conf_thres: float,
                              reg_max: int = 16, max_cand: Optional[int] = None, return_head: bool = False):
    """cvpp_yolov8_head_decode_filter: the head's last 1x1 convolutions + decode + confidence filter in ONE tcgen05
    kernel.  Per level: box_feats[l] (B, c2, H, W), cls_feats[l] (B, c3, H, W) - the inputs of cv2[l][2] / cv3[l][2]
    (modules.py:423-425); box_w[l] (4*reg_max, c2[, 1, 1]), box_b[l] (4*reg_max,), cls_w[l] (nc, c3[, 1, 1]), cls_b[l] (nc,).
    return_head=True also returns the materialised head x_cat (B, 4*reg_max + nc, A) (cvpp_yolov8_head_decode_filter_x)."""
    n = len(box_feats)
    if not (n == len(cls_feats) == len(box_w) == len(box_b) == len(cls_w) == len(cls_b) == len(strides)) or n < 1 or n > 4:
        raise ValueError("between 1 and 4 levels, with one feature pair / weight set / stride each")
    keep = []

    def prep(t, what):
        _require_cuda(t, what)
        t = t.contiguous()
        keep.append(t)
        return t
    bf = [prep(t, f"box_feats[{i}]") for i, t in enumerate(box_feats)]
    cf = [prep(t, f"cls_feats[{i}]") for i, t in enumerate(cls_feats)]
    B, c2, c3 = int(bf[0].shape[0]), int(bf[0].shape[1]), int(cf[0].shape[1])
    bw = [prep(t.reshape(t.shape[0], -1), f"box_w[{i}]") for i, t in enumerate(box_w)]
    cw = [prep(t.reshape(t.shape[0], -1), f"cls_w[{i}]") for i, t in enumerate(cls_w)]
    bb = [prep(t.reshape(-1), f"box_b[{i}]") for i, t in enumerate(box_b)]
    cb = [prep(t.reshape(-1), f"cls_b[{i}]") for i, t in enumerate(cls_b)]
    nc = int(cw[0].shape[0])
    hs, ws = [], []
    dev = bf[0].device
    for i in range(n):
        if bf[i].dim() != 4 or cf[i].dim() != 4 or tuple(bf[i].shape[2:]) != tuple(cf[i].shape[2:]) or \
                bf[i].shape[0] != B or cf[i].shape[0] != B or bf[i].shape[1] != c2 or cf[i].shape[1] != c3:
            raise ValueError(f"level {i}: expected box features (B, {c2}, H, W) and class features (B, {c3}, H, W)")
        if tuple(bw[i].shape) != (4 * reg_max, c2) or tuple(cw[i].shape) != (nc, c3) or bb[i].numel() != 4 * reg_max or \
                cb[i].numel() != nc:
            raise ValueError(f"level {i}: weight / bias shapes do not match the features")
        if any(t.device != dev for t in (bf[i], cf[i], bw[i], cw[i], bb[i], cb[i])):
            raise ValueError("all tensors must live on the same device")
        hs.append(int(bf[i].shape[2]))
        ws.append(int(bf[i].shape[3]))
    A = sum(h * w for h, w in zip(hs, ws))
    max_cand = int(max_cand or A)
    _check_one_key_per_anchor(max_cand, A)
    key = torch.empty((B, max_cand), dtype=torch.int64, device=dev)
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    box_dense = torch.empty((B, A, 4), dtype=torch.float32, device=dev)
    arr = lambda ts: (c_vp * n)(*[t.data_ptr() for t in ts])   # noqa: E731
    head = torch.empty((B, 4 * reg_max + nc, A), dtype=torch.float32, device=dev) if return_head else None
    with torch.cuda.device(dev):
        check(_lib.lib().cvpp_yolov8_head_decode_filter_x(
            arr(bf), arr(cf), arr(bw), arr(bb), arr(cw), arr(cb), (ctypes.c_int * n)(*hs), (ctypes.c_int * n)(*ws),
            (ctypes.c_float * n)(*[float(s) for s in strides]), n, B, c2, c3, nc, int(reg_max), float(conf_thres),
            _ptr(key), _ptr(count), _ptr(box_dense), max_cand, _ptr(head), _stream(dev)))
    cand = Candidates(key, count, box_dense, max_cand, A, nc)
    return (cand, head) if return_head else cand


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
    _check_one_key_per_anchor(max_cand, A)
    dev = pred.device
    key = torch.empty((B, max_cand), dtype=torch.int64, device=dev)
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    box_dense = torch.empty((B, A, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().cvpp_pred_filter(_ptr(pred), B, ch, nc, A, float(conf_thres), _ptr(key), _ptr(count),
                                          _ptr(box_dense), max_cand, _stream(dev)))
    return Candidates(key, count, box_dense, max_cand, A, nc)


def keep_classes(c: Candidates, classes: Sequence[int]) -> None:
    """cvpp_keep_classes: drop (in place) the candidate keys whose class is not in `classes`."""
    ids = [int(v) for v in classes]
    arr = (ctypes.c_int32 * max(len(ids), 1))(*ids)
    dev = c.key.device
    with torch.cuda.device(dev):
        check(_lib.lib().cvpp_keep_classes(_ptr(c.key), _ptr(c.count), int(c.key.shape[0]), c.max_cand, arr, len(ids),
                                           _stream(dev)))


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


def sort_nms(c: Candidates, iou_thres: float, rule: int = RULE_TORCHVISION_CPU, order: int = ORDER_SCORE_DESC,
             max_det: int = 300, max_nms: int = 0, max_out: Optional[int] = None) -> Detections:
    """cvpp_sort_nms: the fused equivalent of segmented_sort(c) + nms(c) on UNSORTED candidate keys."""
    l = _lib.lib()
    B = int(c.key.shape[0])
    dev = c.key.device
    if max_out is None:
        max_out = max_det if (order == ORDER_SCORE_DESC and max_det > 0) else c.max_cand
    max_out = max(int(max_out), 1)
    nbytes = int(l.cvpp_sort_nms_workspace_bytes(B, c.max_cand, c.nc))
    ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=dev)
    det = Detections(box=torch.empty((B, max_out, 4), dtype=torch.float32, device=dev),
                     score=torch.empty((B, max_out), dtype=torch.float32, device=dev),
                     cls=torch.empty((B, max_out), dtype=torch.int32, device=dev),
                     anchor=torch.empty((B, max_out), dtype=torch.int32, device=dev),
                     count=torch.empty((B,), dtype=torch.int32, device=dev), cand_count=c.count)
    with torch.cuda.device(dev):
        check(l.cvpp_sort_nms(_ptr(c.key), _ptr(c.count), _ptr(c.box_dense), B, c.max_cand, c.A, c.nc, float(iou_thres),
                              rule, order, int(max_det), int(max_nms), max_out, _ptr(det.box), _ptr(det.score),
                              _ptr(det.cls), _ptr(det.anchor), _ptr(det.count), _ptr(ws), nbytes, _stream(dev)))
    return det


class Yolov8Postprocessor:
    """Pre-allocated buffers + one C call (cvpp_yolov8_postprocess_ev) per batch: decode+filter, sort, NMS.

    `capture()` records that call into a CUDA graph bound to one LevelSet (fixed input buffers), so a
    steady-state step is a single graph launch - the form to use for bs=1 latency.
    max_cand defaults to A (one key per anchor at most); smaller values are rejected because an overflowing
    candidate buffer would drop keys in a nondeterministic order."""

    def __init__(self, B: int, A: int, nc: int, device, max_det: int = 300, max_cand: Optional[int] = None):
        self.B, self.A, self.nc, self.max_det = int(B), int(A), int(nc), int(max_det)
        self.max_cand = int(max_cand or A)
        if self.max_cand < self.A:
            raise ValueError(f"max_cand={self.max_cand} < A={self.A}: the candidate buffer could overflow "
                             "(every anchor can pass the confidence filter)")
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
                 max_nms: int = 30000, reg_max: int = 16, consumed: Optional[torch.cuda.Event] = None,
                 gather: Optional[Tuple[Sequence[int], int, int]] = None) -> Detections:
        """`consumed` (optional CUDA event) is recorded right after the decode kernel, the last reader of `ls`.
        `gather` = (peer_ptrs, multicast_ptr, rank): the NMS kernel also writes the image's CVPP_ROWS_FULL rows + count into
        every rank's gather buffer (cvpp_yolov8_postprocess_gather: the all-gather without a third launch)."""
        if ls.B != self.B or ls.A != self.A or ls.C != 4 * reg_max + self.nc:
            raise ValueError("level set does not match the shapes this post-processor was built for")
        dev = self.device
        if ls.device != dev:
            raise ValueError(f"level set lives on {ls.device}, this post-processor on {dev}")
        if torch.cuda.current_device() != dev.index:
            with torch.cuda.device(dev):
                return self.__call__(ls, conf_thres, iou_thres, rule, max_nms, reg_max, consumed, gather)
        ev = c_vp(0)
        if consumed is not None:
            if not consumed.cuda_event:          # torch creates the handle lazily, on the first record
                consumed.record()
            ev = c_vp(consumed.cuda_event)
        if gather is not None:
            peer_ptrs, mc_ptr, rank = gather
            n = len(peer_ptrs)
            arr = (c_vp * n)(*[int(q) for q in peer_ptrs])
            check(self._lib.cvpp_yolov8_postprocess_gather(
                ls.ptr, ls.batch_stride, ls.chan_stride, ls.h, ls.w, ls.stride, ls.n, self.B, self.nc, reg_max,
                float(conf_thres), float(iou_thres), rule, self.max_det, int(max_nms), self.max_cand, *self._out_args,
                ev, arr, c_vp(int(mc_ptr or 0)), n, int(rank), c_vp(torch.cuda.current_stream().cuda_stream)))
            return self.det
        check(self._lib.cvpp_yolov8_postprocess_ev(
            ls.ptr, ls.batch_stride, ls.chan_stride, ls.h, ls.w, ls.stride, ls.n, self.B, self.nc, reg_max,
            float(conf_thres), float(iou_thres), rule, self.max_det, int(max_nms), self.max_cand, *self._out_args,
            ev, c_vp(torch.cuda.current_stream().cuda_stream)))
        return self.det

    def capture(self, ls: LevelSet, conf_thres: float, iou_thres: float, rule: int = RULE_TORCHVISION_CPU,
                max_nms: int = 30000, reg_max: int = 16, consumed: Optional[torch.cuda.Event] = None) -> "GraphedPostprocess":
        return GraphedPostprocess(self, ls, (conf_thres, iou_thres, rule, max_nms, reg_max, consumed))


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


class PipelinedPostprocess:
    """Throughput mode of the YOLOv8 post-processor: `depth` slots, each with its own INPUT buffers (one LevelSet per
    slot), detection buffers, CUDA graph and stream, used round-robin.  Consecutive batches are independent, so the
    fused sort+NMS kernel of batch k (64 CTAs, latency-bound) overlaps the HBM-bound decode of batch k+1 on the SMs
    it leaves idle.

    Synchronisation with the producer of the head tensors (the network, or an H2D copy):
      * `submit(ready)` makes the slot's stream wait for `ready` (an event the producer recorded after writing the
        slot's inputs; None = the inputs were written on the current stream before fork()/submit order holds);
      * `consumed[slot]` is recorded right after the slot's decode kernel (inside the replayed graph, as an external
        event-record node): the producer waits on it before overwriting that slot's inputs - while the slot's NMS
        may still be running;
      * `wait(slot)` / `join()`: the caller's stream waits for the slot's detections / for everything.

        pipe = PipelinedPostprocess(B, A, nc, device, [ls0, ls1, ls2], conf, iou)
        for k, batch in enumerate(batches):
            slot = pipe.next_slot
            with torch.cuda.stream(producer):
                producer.wait_event(pipe.consumed[slot])     # slot's previous decode is done with the buffers
                write(batch, into=inputs[slot]); ready = torch.cuda.Event(); ready.record(producer)
            det = pipe.submit(ready)                          # returns at once
        pipe.join()

    A single LevelSet may be passed instead of a list: every slot then reads the SAME buffers (static-input
    benchmarking only - a real producer could not refill them while another slot's decode is in flight).

    Multi-GPU evaluation: `gather=distributed.PeerGather(...)` (at least `depth` slots) makes `submit(gather=True)` also store
    the slot's CVPP_ROWS_FULL rows + counts into gather slot `slot` of every rank, from inside the slot's CUDA graph - written
    by the NMS kernel's own CTAs (`fused_rows=True`, cvpp_yolov8_postprocess_gather: no third launch) or by the epilogue
    kernel behind a programmatic dependent launch; `use_multicast` picks the NVSwitch multicast address when the gather
    has one.  Fence with `gather.barrier(slot)` before reading `gather.view(slot)`."""

    def __init__(self, B: int, A: int, nc: int, device, inputs, conf_thres: float, iou_thres: float,
                 max_det: int = 300, depth: Optional[int] = None, graph: bool = True, gather=None,
                 use_multicast: bool = True, fused_rows: bool = True, **post_args):
        self.use_multicast = bool(use_multicast)
        self.fused_rows = bool(fused_rows)
        if isinstance(inputs, LevelSet):
            depth = int(depth or 2)
            inputs = [inputs] * depth
            self.static_input = True
        else:
            inputs = list(inputs)
            if depth is not None and int(depth) != len(inputs):
                raise ValueError(f"depth={depth} but {len(inputs)} input level sets were given (one per slot)")
            depth = len(inputs)
            self.static_input = False
        if depth < 1:
            raise ValueError("at least one slot is needed")
        self.inputs = inputs
        self.posts = [Yolov8Postprocessor(B, A, nc, device, max_det=max_det) for _ in range(depth)]
        self.device = self.posts[0].device
        self.args = (conf_thres, iou_thres) + tuple(post_args.get(k, d) for k, d in
                                                    (("rule", RULE_TORCHVISION_CPU), ("max_nms", 30000), ("reg_max", 16)))
        self.consumed = [torch.cuda.Event() for _ in range(depth)]
        with torch.cuda.device(self.device):
            for e in self.consumed:
                e.record()                  # creates the handle; "consumed" holds before the first use of a slot
        self.graphs = [pp.capture(ls, *self.args, consumed=e) for pp, ls, e in zip(self.posts, inputs, self.consumed)] \
            if graph else None
        # optional: the slot's graph also holds the fused epilogue + all-gather of the slot's detections into gather slot
        # `slot` of a distributed.PeerGather (a programmatic dependent launch behind the NMS kernel: no launch gap, and one
        # graph launch per step on the host instead of a graph launch plus an eager kernel)
        self.gather = gather
        self.graphs_gather = None
        if gather is not None:
            if gather.depth < depth:
                raise ValueError(f"the PeerGather has {gather.depth} slots, the pipeline {depth}")
            if graph:
                self.graphs_gather = []
                with torch.cuda.device(self.device):
                    for i, (pp, ls, e) in enumerate(zip(self.posts, inputs, self.consumed)):
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            self._post_and_gather(i)
                        self.graphs_gather.append(g)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(depth)]
        self.turn = 0
        self.last_slot = 0

    @property
    def ls(self) -> LevelSet:
        return self.inputs[0]

    @property
    def next_slot(self) -> int:
        return self.turn % len(self.posts)

    def fork(self) -> None:
        """Every pipeline stream waits for the work already queued on the caller's current stream."""
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)

    def _post_and_gather(self, i: int) -> None:
        mc = self.gather.multicast_ptr(i) if self.use_multicast else 0
        if self.fused_rows:   # the NMS kernel writes the rows itself: two launches per step
            self.posts[i](self.inputs[i], *self.args, consumed=self.consumed[i], gather=(self.gather.peer_ptrs(i), mc, self.gather.rank))
            return
        det = self.posts[i](self.inputs[i], *self.args, consumed=self.consumed[i])
        detection_epilogue_allgather(det, ROWS_FULL, self.gather.peer_ptrs(i), self.gather.rank, multicast_ptr=mc)

    def submit(self, ready: Optional[torch.cuda.Event] = None, gather: bool = False) -> Detections:
        """gather=True (needs the `gather=` PeerGather of the constructor): the slot's rows + counts are also stored into
        gather slot `slot` of every rank; fence with gather.barrier(slot) before reading gather.view(slot)."""
        i = self.turn % len(self.posts)
        self.turn += 1
        self.last_slot = i
        st = self.streams[i]
        if ready is not None:
            st.wait_event(ready)
        if gather and self.gather is None:
            raise ValueError("submit(gather=True) needs PipelinedPostprocess(..., gather=PeerGather)")
        with torch.cuda.stream(st):
            if gather:
                if self.graphs_gather is not None:
                    self.graphs_gather[i].replay()
                else:
                    self._post_and_gather(i)
            elif self.graphs is not None:
                self.graphs[i].replay()
            else:
                self.posts[i](self.inputs[i], *self.args, consumed=self.consumed[i])
        return self.posts[i].det

    def wait(self, slot: int) -> None:
        torch.cuda.current_stream(self.device).wait_stream(self.streams[slot])

    def join(self) -> None:
        for i in range(len(self.streams)):
            self.wait(i)


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


def _anchor_table(anchors, n_levels: int) -> ctypes.Array:
    flat = [float(v) for row in anchors for v in (row if hasattr(row, "__len__") else [row])]
    if len(flat) != n_levels * 6:
        raise ValueError(f"expected {n_levels * 3} (w, h) anchors, 3 per level, got {len(flat) // 2}")
    return (ctypes.c_float * len(flat))(*flat)


def yolov7_decode_filter(ls: LevelSet, nc: int, anchors, input_hw: Sequence[int], conf_thres: float,
                         max_cand: Optional[int] = None) -> Candidates:
    """ls: levels (B, 3*(5+nc), H, W) in the reference's order; anchors: (3*levels, 2) pixels, row 3*l + a =
    anchors[anchors_mask[l]][a].  Candidate keys + normalised xyxy + (obj, class_conf) per anchor."""
    if ls.C != 3 * (5 + nc):
        raise ValueError(f"head has {ls.C} channels, expected 3*(5+{nc})")
    A = 3 * ls.A
    max_cand = int(max_cand or A)
    dev = ls.device
    key = torch.empty((ls.B, max_cand), dtype=torch.int64, device=dev)
    count = torch.empty((ls.B,), dtype=torch.int32, device=dev)
    box_dense = torch.empty((ls.B, A, 4), dtype=torch.float32, device=dev)
    aux_dense = torch.empty((ls.B, A, 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().cvpp_yolov7_decode_filter(ls.ptr, ls.batch_stride, ls.chan_stride, ls.h, ls.w,
                                                   _anchor_table(anchors, ls.n), ls.n, ls.B, nc, int(input_hw[0]),
                                                   int(input_hw[1]), float(conf_thres), _ptr(key), _ptr(count),
                                                   _ptr(box_dense), _ptr(aux_dense), max_cand, _stream(dev)))
    return Candidates(key, count, box_dense, max_cand, A, nc, aux_dense)


def yolov7_pred_filter(pred: torch.Tensor, nc: int, conf_thres: float, max_cand: Optional[int] = None) -> Candidates:
    """pred (B, A, 5+nc) decoded (cx, cy, w, h, obj, cls...) -> candidates (YOLOv7._nms / yolo7_nms front half)."""
    _require_cuda(pred, "prediction")
    if pred.dim() != 3 or pred.shape[2] < 5 + nc:
        raise ValueError(f"prediction must be (B, A, 5+nc), got {tuple(pred.shape)}")
    if pred.shape[2] != 5 + nc:
        pred = pred[..., :5 + nc]
    pred = pred.contiguous()
    B, A = int(pred.shape[0]), int(pred.shape[1])
    max_cand = int(max_cand or A)
    dev = pred.device
    key = torch.empty((B, max_cand), dtype=torch.int64, device=dev)
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    box_dense = torch.empty((B, A, 4), dtype=torch.float32, device=dev)
    aux_dense = torch.empty((B, A, 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().cvpp_yolov7_pred_filter(_ptr(pred), B, A, nc, float(conf_thres), _ptr(key), _ptr(count),
                                                 _ptr(box_dense), _ptr(aux_dense), max_cand, _stream(dev)))
    return Candidates(key, count, box_dense, max_cand, A, nc, aux_dense)


def yolov3_decode_filter(ls: LevelSet, nc: int, anchors, input_hw: Sequence[int], conf_thres: float,
                         merge_batch: bool = False, max_cand: Optional[int] = None) -> Candidates:
    """ls: levels (B, 3*(5+nc), H, W) (13, 26, 52 for 416^2); anchors (3*levels, 2) pixels in level order.
    One key per (anchor, class) with sigmoid(obj)*sigmoid(cls) >= conf.  merge_batch flattens the batch
    into one output image like the reference Decoder."""
    if ls.C != 3 * (5 + nc):
        raise ValueError(f"head has {ls.C} channels, expected 3*(5+{nc})")
    A = 3 * ls.A * (ls.B if merge_batch else 1)
    n_out = (1 if ls.B > 0 else 0) if merge_batch else ls.B
    max_cand = int(max_cand or min(A * nc, 1 << 16))
    dev = ls.device
    key = torch.empty((n_out, max_cand), dtype=torch.int64, device=dev)
    count = torch.empty((n_out,), dtype=torch.int32, device=dev)
    box_dense = torch.empty((n_out, A, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().cvpp_yolov3_decode_filter(ls.ptr, ls.batch_stride, ls.chan_stride, ls.h, ls.w,
                                                   _anchor_table(anchors, ls.n), ls.n, ls.B, nc, int(input_hw[0]),
                                                   int(input_hw[1]), float(conf_thres), int(bool(merge_batch)),
                                                   _ptr(key), _ptr(count), _ptr(box_dense), max_cand, _stream(dev)))
    return Candidates(key, count, box_dense, max_cand, A, nc)


def yolov3_predict_bbox(feature: torch.Tensor, nc: int, anchors_norm):
    """predict_bounding_bbox: feature (N, 3*(5+nc), H, W), anchors (3, 2) normalised ->
    (box_xy (N,H,W,3,2), box_wh (N,H,W,3,2), confidence (N,H,W,3,1), class_prob (N,H,W,3,nc))."""
    _require_cuda(feature, "feature_map")
    feature = feature.contiguous()
    N, C, H, W = (int(v) for v in feature.shape)
    if C != 3 * (5 + nc):
        raise ValueError(f"feature map has {C} channels, expected 3*(5+{nc})")
    dev = feature.device
    xy = torch.empty((N, H, W, 3, 2), dtype=torch.float32, device=dev)
    wh = torch.empty((N, H, W, 3, 2), dtype=torch.float32, device=dev)
    conf = torch.empty((N, H, W, 3, 1), dtype=torch.float32, device=dev)
    prob = torch.empty((N, H, W, 3, nc), dtype=torch.float32, device=dev)
    flat = [float(v) for row in anchors_norm for v in row]
    if len(flat) != 6:
        raise ValueError("expected 3 (w, h) anchors")
    with torch.cuda.device(dev):
        check(_lib.lib().cvpp_yolov3_predict_bbox(_ptr(feature), N, nc, H, W, (ctypes.c_float * 6)(*flat), _ptr(xy),
                                                  _ptr(wh), _ptr(conf), _ptr(prob), _stream(dev)))
    return xy, wh, conf, prob


def score_matrix_filter(boxes: torch.Tensor, scores: torch.Tensor, conf_thres: float,
                        max_cand: Optional[int] = None) -> Candidates:
    """boxes (M, 4) xyxy + scores (M, nc) -> single-image candidates, one key per score >= conf."""
    _require_cuda(boxes, "boxes")
    _require_cuda(scores, "scores")
    boxes, scores = boxes.contiguous(), scores.contiguous()
    M, nc = int(scores.shape[0]), int(scores.shape[1])
    if tuple(boxes.shape) != (M, 4):
        raise ValueError("expected boxes (M, 4) and scores (M, nc)")
    max_cand = int(max_cand or max(min(M * nc, 1 << 16), 1))
    if boxes.data_ptr() % 16:
        boxes = boxes.clone()
    dev = boxes.device
    while True:
        key = torch.empty((1, max_cand), dtype=torch.int64, device=dev)
        count = torch.empty((1,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().cvpp_score_matrix_filter(_ptr(scores), M, nc, float(conf_thres), _ptr(key), _ptr(count),
                                                      max_cand, _stream(dev)))
        n = int(count.item())
        if n <= max_cand:
            return Candidates(key, count, boxes.reshape(1, M, 4), max_cand, M, nc)
        max_cand = n


def gather_feat(feat: torch.Tensor, ind: torch.Tensor, count: Optional[torch.Tensor] = None,
                check_bounds: bool = True) -> torch.Tensor:
    """feat (B, N, C) float32, ind (B, K) int32/int64 -> (B, K, C) rows feat[b, ind[b, k]]."""
    _require_cuda(feat, "feat")
    if ind.dtype not in (torch.int32, torch.int64) or not ind.is_cuda:
        raise ValueError("ind must be an int32 / int64 CUDA tensor")
    feat, ind = feat.contiguous(), ind.contiguous()
    B, N, C = (int(v) for v in feat.shape)
    K = int(ind.shape[1])
    dev = feat.device
    out = torch.zeros((B, K, C), dtype=torch.float32, device=dev) if count is not None else \
        torch.empty((B, K, C), dtype=torch.float32, device=dev)
    err = torch.empty((1,), dtype=torch.int32, device=dev) if check_bounds else None
    with torch.cuda.device(dev):
        check(_lib.lib().cvpp_gather_feat(_ptr(feat), _ptr(ind), int(ind.dtype == torch.int64), _ptr(count), B, N, C, K,
                                          _ptr(out), _ptr(err), _stream(dev)))
    if check_bounds and int(err.item()):
        raise IndexError("gather_feat: index out of range")
    return out


def detection_epilogue(det: Detections, layout: int, box_mode: int = BOX_KEEP, letterbox: Optional[torch.Tensor] = None,
                       aux_dense: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                       packed: bool = False) -> torch.Tensor:
    """cvpp_detection_epilogue: (B, max_out, 6|7) caller-facing rows on the device.
    packed=True returns ONE flat fp32 buffer [B*max_out*W rows | B counts] (the all-gather payload; see
    distributed.gather_packed / unpack_detections); `out` reuses a buffer of the right size."""
    B, max_out = int(det.box.shape[0]), int(det.box.shape[1])
    dev = det.box.device
    width = 7 if layout in (ROWS_YOLOV7, ROWS_FULL) else 6
    n_rows = B * max_out * width
    if out is None:
        out = torch.empty((n_rows + (B if packed else 0),), dtype=torch.float32, device=dev)
    flat = out.reshape(-1)
    if flat.numel() != n_rows + (B if packed else 0) or flat.dtype != torch.float32 or not flat.is_contiguous():
        raise ValueError("`out` has the wrong size / dtype for this epilogue")
    count_ptr = c_vp(flat.data_ptr() + 4 * n_rows) if packed else c_vp(0)
    A = int(aux_dense.shape[1]) if aux_dense is not None else 0
    if letterbox is not None:
        letterbox = letterbox.to(device=dev, dtype=torch.float32).contiguous()
        if tuple(letterbox.shape) != (B, 5):
            raise ValueError("letterbox must be (B, 5)")
    with torch.cuda.device(dev):
        check(_lib.lib().cvpp_detection_epilogue(_ptr(det.box), _ptr(det.score), _ptr(det.cls), _ptr(det.anchor),
                                                 _ptr(det.count), _ptr(aux_dense), B, max_out, A, int(layout),
                                                 int(box_mode), _ptr(letterbox), _ptr(flat), count_ptr, _stream(dev)))
    return flat if packed else flat.reshape(B, max_out, width)


def detection_epilogue_compact(det: Detections, layout: int, row_capacity: int, box_mode: int = BOX_KEEP,
                               letterbox: Optional[torch.Tensor] = None, aux_dense: Optional[torch.Tensor] = None,
                               out: Optional[torch.Tensor] = None, row_offset: Optional[torch.Tensor] = None,
                               overflow: Optional[torch.Tensor] = None):
    """cvpp_detection_epilogue_compact: rows of all images back to back (no padding).  Returns
    (rows (row_capacity, W) float32, row_offset (B + 1,) int32, overflow (1,) int32) - all on the device, nothing is
    read back; image b owns rows[row_offset[b]:row_offset[b + 1]]."""
    B, max_out = int(det.box.shape[0]), int(det.box.shape[1])
    dev = det.box.device
    width = 7 if layout in (ROWS_YOLOV7, ROWS_FULL) else 6
    row_capacity = int(row_capacity)
    if out is None:
        out = torch.empty((max(row_capacity, 1), width), dtype=torch.float32, device=dev)
    if out.numel() < row_capacity * width or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("`out` is too small / not contiguous float32 for this row capacity")
    if row_offset is None:
        row_offset = torch.empty((B + 1,), dtype=torch.int32, device=dev)
    if overflow is None:
        overflow = torch.empty((1,), dtype=torch.int32, device=dev)
    A = int(aux_dense.shape[1]) if aux_dense is not None else 0
    if letterbox is not None:
        letterbox = letterbox.to(device=dev, dtype=torch.float32).contiguous()
        if tuple(letterbox.shape) != (B, 5):
            raise ValueError("letterbox must be (B, 5)")
    with torch.cuda.device(dev):
        check(_lib.lib().cvpp_detection_epilogue_compact(_ptr(det.box), _ptr(det.score), _ptr(det.cls), _ptr(det.anchor),
                                                         _ptr(det.count), _ptr(aux_dense), B, max_out, A, int(layout),
                                                         int(box_mode), _ptr(letterbox), _ptr(out), row_capacity,
                                                         _ptr(row_offset), _ptr(overflow), _stream(dev)))
    return out, row_offset, overflow


def detection_epilogue_allgather(det: Detections, layout: int, peer_ptrs: Sequence[int], rank: int,
                                 box_mode: int = BOX_KEEP, letterbox: Optional[torch.Tensor] = None,
                                 aux_dense: Optional[torch.Tensor] = None, multicast_ptr: int = 0) -> None:
    """cvpp_detection_epilogue_allgather: rows + counts stored into every rank's gather buffer (peer-mapped
    device pointers `peer_ptrs`, one per rank) at slot `rank`.  With `multicast_ptr` (the NVSwitch multicast address of
    the same buffer, PeerGather.multicast_ptr(i)) the rows leave the GPU once as multimem.st
    (cvpp_detection_epilogue_multicast).  See distributed.PeerGather."""
    B, max_out = int(det.box.shape[0]), int(det.box.shape[1])
    dev = det.box.device
    A = int(aux_dense.shape[1]) if aux_dense is not None else 0
    n = len(peer_ptrs)
    if letterbox is not None:
        letterbox = letterbox.to(device=dev, dtype=torch.float32).contiguous()
    if multicast_ptr:
        with torch.cuda.device(dev):
            check(_lib.lib().cvpp_detection_epilogue_multicast(_ptr(det.box), _ptr(det.score), _ptr(det.cls), _ptr(det.anchor),
                                                               _ptr(det.count), _ptr(aux_dense), B, max_out, A, int(layout),
                                                               int(box_mode), _ptr(letterbox), ctypes.c_void_p(int(multicast_ptr)),
                                                               n, int(rank), _stream(dev)))
        return
    arr = (c_vp * n)(*[int(p) for p in peer_ptrs])
    with torch.cuda.device(dev):
        check(_lib.lib().cvpp_detection_epilogue_allgather(_ptr(det.box), _ptr(det.score), _ptr(det.cls), _ptr(det.anchor),
                                                           _ptr(det.count), _ptr(aux_dense), B, max_out, A, int(layout),
                                                           int(box_mode), _ptr(letterbox), arr, n, int(rank), _stream(dev)))


def letterbox_reverse(boxes: torch.Tensor, h: float, w: float, input_size: Sequence[int], xywh: bool = True) -> torch.Tensor:
    """cvpp_letterbox_reverse: reverse_letter_box on CUDA boxes (..., 4) -> (..., 4) corners in original pixels."""
    _require_cuda(boxes, "boxes")
    shape = boxes.shape
    flat = boxes.reshape(-1, 4).contiguous()
    if flat.data_ptr() % 16:
        flat = flat.clone()
    out = torch.empty_like(flat)
    scale = max(h / input_size[0], w / input_size[1])        # Python doubles, like the reference (:115-119)
    top = (input_size[0] - h / scale) // 2
    left = (input_size[1] - w / scale) // 2
    with torch.cuda.device(flat.device):
        check(_lib.lib().cvpp_letterbox_reverse(_ptr(flat), int(flat.shape[0]), int(bool(xywh)), float(input_size[1]),
                                                float(input_size[0]), float(left), float(top), float(scale), _ptr(out),
                                                _stream(flat.device)))
    return out.reshape(shape)


def centernet_suppress(heat: torch.Tensor) -> torch.Tensor:
    """cvpp_centernet_suppress: heat (B, H, W, C) -> heat * (heat == maxpool3x3 over (x, channel))."""
    _require_cuda(heat, "heatmap")
    if heat.dim() != 4:
        raise ValueError(f"heatmap must be (B, H, W, C), got {tuple(heat.shape)}")
    heat = heat.contiguous()
    B, H, W, C = (int(v) for v in heat.shape)
    out = torch.empty_like(heat)
    with torch.cuda.device(heat.device):
        check(_lib.lib().cvpp_centernet_suppress(_ptr(heat), B, H, W, C, _ptr(out), _stream(heat.device)))
    return out


def topk(scores: torch.Tensor, k: int, split: Optional[Tuple[int, int]] = None):
    """cvpp_topk: scores (B, N) float32 -> (values (B, k), indices (B, k) int64), scores descending, equal scores by
    the lower flat index.  split=(C, W) also returns the reference's index split (CenterNetA._top_k,
    centernet.py:331-337): (values, indices, cls, ys, xs, pixel) with pixel = y*W + x as int32."""
    _require_cuda(scores, "scores")
    if scores.dim() != 2:
        raise ValueError(f"scores must be (B, N), got {tuple(scores.shape)}")
    scores = scores.contiguous()
    B, N = int(scores.shape[0]), int(scores.shape[1])
    k = int(k)
    if k < 1 or k > N:
        raise RuntimeError("selected index k out of range")      # torch.topk's message for the same misuse
    dev = scores.device
    l = _lib.lib()
    nbytes = int(l.cvpp_topk_workspace_bytes(B, k))
    ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=dev)
    val = torch.empty((B, k), dtype=torch.float32, device=dev)
    idx = torch.empty((B, k), dtype=torch.int64, device=dev)
    cls = ys = xs = pix = None
    C = W = 0
    if split is not None:
        C, W = int(split[0]), int(split[1])
        cls, ys, xs = (torch.empty((B, k), dtype=torch.int64, device=dev) for _ in range(3))
        pix = torch.empty((B, k), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(l.cvpp_topk(_ptr(scores), B, N, k, C, W, _ptr(val), _ptr(idx), _ptr(cls), _ptr(ys), _ptr(xs), _ptr(pix),
                          _ptr(ws), nbytes, _stream(dev)))
    return (val, idx) if split is None else (val, idx, cls, ys, xs, pix)


def voc_match(det_rows: torch.Tensor, det_offset: torch.Tensor, gt_box: torch.Tensor, gt_cls: torch.Tensor,
              gt_difficult: torch.Tensor, gt_offset: torch.Tensor, min_overlap: float):
    """cvpp_voc_match: det_rows (N, 6) float32 VOC rows, det_offset (B+1,) int32, gt_box (G, 4) float32, gt_cls /
    gt_difficult (G,) int32, gt_offset (B+1,) int32 - all on the device -> (flag (N,) int32: 1 TP / 2 FP / 0 difficult
    match, best_gt (N,) int32, ovmax (N,) float64)."""
    _require_cuda(det_rows, "det_rows")
    _require_cuda(gt_box, "gt_box")
    dev = det_rows.device
    for t, name in ((det_offset, "det_offset"), (gt_cls, "gt_cls"), (gt_difficult, "gt_difficult"), (gt_offset, "gt_offset")):
        if t.dtype != torch.int32 or t.device != dev:
            raise ValueError(f"{name} must be an int32 tensor on {dev}")
    det_rows, gt_box = det_rows.contiguous(), gt_box.contiguous()
    B = int(det_offset.numel()) - 1
    if B < 0 or int(gt_offset.numel()) != B + 1 or det_rows.dim() != 2 or det_rows.shape[1] != 6:
        raise ValueError("expected det_rows (N, 6) and offsets with B + 1 entries each")
    N, G = int(det_rows.shape[0]), int(gt_box.shape[0])
    flag = torch.empty((max(N, 1),), dtype=torch.int32, device=dev)
    best = torch.empty((max(N, 1),), dtype=torch.int32, device=dev)
    ov = torch.empty((max(N, 1),), dtype=torch.float64, device=dev)
    claim = torch.empty((max(G, 1),), dtype=torch.int32, device=dev)
    if gt_box.numel() == 0:
        gt_box = torch.zeros((1, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().cvpp_voc_match(_ptr(det_rows), _ptr(det_offset.contiguous()), _ptr(gt_box), _ptr(gt_cls.contiguous()),
                                        _ptr(gt_difficult.contiguous()), _ptr(gt_offset.contiguous()), B, float(min_overlap),
                                        _ptr(flag), _ptr(best), _ptr(ov), _ptr(claim), _stream(dev)))
    return flag[:N], best[:N], ov[:N]


def correct_boxes_params(image_hw: Sequence[Tuple[int, int]], input_hw: Sequence[int], letterbox_image: bool,
                         device) -> torch.Tensor:
    """The (B, 5) table cvpp_detection_epilogue wants for yolo_correct_boxes (image_process.py:161-181)."""
    if letterbox_image:
        return letterbox_params(image_hw, input_hw, device)
    rows = [[float(w), float(h), 0.0, 0.0, 1.0] for (h, w) in image_hw]
    return torch.tensor(rows, dtype=torch.float32).to(device)


def per_class_nms_device(c: Candidates, nms_thres: float, initial_out: int = 4096) -> Detections:
    """sort + per-class NMS (class-major order, no cap), result left on the device; retries with a larger
    output capacity when the first guess overflows (one scalar D2H read of the max count)."""
    if int(c.count.max().item()) > c.max_cand:
        raise OverflowError("candidate buffer overflow: re-run the filter with a larger max_cand")
    max_out = max(min(c.max_cand, initial_out), 1)
    while True:
        det = sort_nms(c, nms_thres, RULE_PER_CLASS, ORDER_CLASS_MAJOR, max_det=0, max_out=max_out)
        need = int(det.count.max().item()) if det.count.numel() else 0
        if need <= max_out:
            return det
        max_out = need


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
