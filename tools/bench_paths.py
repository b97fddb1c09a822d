"""Per-configuration device timings of every decode path (BASELINE.json configs[1..4] + YOLOv3), with the
algorithmic-bytes bandwidth of each decode/filter kernel.  One JSON line per configuration.

    python tools/bench_paths.py [--iters 50] [--only yolov8,centernet,ssd,yolov7,yolov3]

Inputs are generated on the device with the SURVEY.md §8d distributions (no score separation: this is a
timing tool, parity lives in tests/)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from computervision.pytorch_b200 import ops  # noqa: E402
from computervision.pytorch_b200.core.utils.anchor import generate_ssd_anchor_v2  # noqa: E402

# the reference's config constants (configs/ssd_cfg.py, yolo7_cfg.py, yolo3 cfg): SSD300 prior recipe, YOLOv7 anchors in
# the order of anchors_mask ((6,7,8),(3,4,5),(0,1,2)), YOLOv3 anchors in level order
SSD_SIZES, SSD_FEATS = (30, 60, 111, 162, 213, 264, 315), (38, 19, 10, 5, 3, 1)
SSD_RATIOS = ((1, 2, 0.5), (1, 2, 0.5, 3, 1.0 / 3), (1, 2, 0.5, 3, 1.0 / 3), (1, 2, 0.5, 3, 1.0 / 3), (1, 2, 0.5), (1, 2, 0.5))
YOLOV7_LEVEL_ANCHORS = np.array([142, 110, 192, 243, 459, 401, 36, 75, 76, 55, 72, 146, 12, 16, 19, 36, 40, 28],
                                np.float32).reshape(-1, 2)
YOLOV3_LEVEL_ANCHORS = np.array([116, 90, 156, 198, 373, 326, 30, 61, 62, 45, 59, 119, 10, 13, 16, 30, 33, 23],
                                np.float32).reshape(-1, 2)

DEV = torch.device("cuda:0")
PEAK = 6504.1
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def timed_graph(fn, iters, warm=3):
    """fn captured once into a CUDA graph (its allocations live in the graph's pool) and replayed: the device time of
    short kernels without the Python / ctypes issue time of the eager call."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return timed(g.replay, iters, warm)


def report(name, B, bytes_per_image, ms_decode, ms_total, extra=None):
    gbs = B * bytes_per_image / (ms_decode * 1e-3) / 1e9
    line = {"path": name, "batch": B, "decode_ms": round(ms_decode, 4), "total_ms": round(ms_total, 4),
            "decode_GBps": round(gbs, 1), "decode_frac_of_hbm_peak": round(gbs / PEAK, 4),
            "images_per_s": round(B / (ms_total * 1e-3), 1), "bytes_per_image": bytes_per_image}
    line.update(extra or {})
    return line


def gen():
    g = torch.Generator(device=DEV)
    g.manual_seed(99)
    return g


def bench_yolov8(iters, B=64):
    g = gen()
    levels = []
    for h, w in ((80, 80), (40, 40), (20, 20)):
        x = torch.randn((B, 144, h, w), generator=g, device=DEV)
        x[:, :64] *= 3.0
        x[:, 64:] *= 4.3155
        x[:, 64:] += -18.19
        levels.append(x)
    ls = ops.make_levels(levels, (8.0, 16.0, 32.0))
    post = ops.Yolov8Postprocessor(B, 8400, 80, DEV)
    c = ops.yolov8_decode_filter(ls, 80, 0.001)
    ms_dec = timed(lambda: ops.yolov8_decode_filter(ls, 80, 0.001), iters)
    ms_tot = timed(lambda: post(ls, 0.001, 0.7), iters)
    cand = float(c.count.float().mean())
    return report("yolov8_C2", B, 144 * 8400 * 4 + 24 * cand, ms_dec, ms_tot, {"cand_per_image": cand})


def bench_centernet(iters, B=64):
    g = gen()
    pred = torch.empty((B, 128, 128, 84), device=DEV)
    pred[..., :80] = torch.randn((B, 128, 128, 80), generator=g, device=DEV) * 1.5 - 5.0
    pred[..., 80:82] = torch.rand((B, 128, 128, 2), generator=g, device=DEV)
    pred[..., 82:] = torch.rand((B, 128, 128, 2), generator=g, device=DEV) * 20
    ms = timed(lambda: ops.centernet_decode(pred, 100, 0.001), iters)
    ms_nms = timed(lambda: ops.centernet_decode(pred, 100, 0.001, use_nms=True), iters)
    return report("centernet_C3", B, 128 * 128 * 84 * 4, ms, ms, {"total_ms_with_diou_nms": round(ms_nms, 4)})


def bench_ssd(iters, B=128):
    g = gen()
    P = 8732
    loc = torch.randn((B, P, 4), generator=g, device=DEV)
    conf = torch.randn((B, P, 21), generator=g, device=DEV) * 2.0
    conf[..., 0] = torch.randn((B, P), generator=g, device=DEV) + 12.0
    boost = torch.rand((B, P), generator=g, device=DEV) < 0.003
    cls = torch.randint(1, 21, (B, P), generator=g, device=DEV)
    val = conf[..., 0] + torch.randn((B, P), generator=g, device=DEV) * 2.0 + 2.0
    conf.scatter_(2, cls.unsqueeze(2), torch.where(boost, val, conf.gather(2, cls.unsqueeze(2)).squeeze(2)).unsqueeze(2))
    pri = torch.from_numpy(generate_ssd_anchor_v2((300, 300), SSD_SIZES, SSD_FEATS, SSD_RATIOS)).to(DEV)
    c = ops.ssd_decode_filter(loc, conf, pri, 0.001, max_cand=32768)
    ms_dec = timed(lambda: ops.ssd_decode_filter(loc, conf, pri, 0.001, max_cand=32768), iters)

    def full():
        cc = ops.ssd_decode_filter(loc, conf, pri, 0.001, max_cand=32768)
        ops.sort_nms(cc, 0.5, ops.RULE_PER_CLASS, ops.ORDER_CLASS_MAJOR, max_det=0, max_out=4096)
    ms_tot = timed(full, iters)
    return report("ssd_C4", B, 873200, ms_dec, ms_tot, {"cand_per_image": float(c.count.float().mean())})


def yolov7_inputs(B, g):
    levels = []
    for s in (20, 40, 80):
        x = torch.randn((B, 3, 85, s, s), generator=g, device=DEV)
        x[:, :, 4] = x[:, :, 4] * 3.0 - 9.0
        x[:, :, 5:] = x[:, :, 5:] * 2.0 - 1.0
        levels.append(x.reshape(B, 255, s, s))
    return levels


def bench_yolov7(iters, B=128):
    levels = yolov7_inputs(B, gen())
    ls = ops.make_levels(levels)
    anchors = YOLOV7_LEVEL_ANCHORS
    c = ops.yolov7_decode_filter(ls, 80, anchors, (640, 640), 0.001)
    ms_dec = timed(lambda: ops.yolov7_decode_filter(ls, 80, anchors, (640, 640), 0.001), iters)

    def full():
        cc = ops.yolov7_decode_filter(ls, 80, anchors, (640, 640), 0.001)
        ops.sort_nms(cc, 0.3, ops.RULE_PER_CLASS, ops.ORDER_CLASS_MAJOR, max_det=0, max_out=8192)
    ms_tot = timed(full, iters)
    return report("yolov7_C5_shard", B, 25200 * 85 * 4, ms_dec, ms_tot, {"cand_per_image": float(c.count.float().mean())})


def bench_yolov3(iters, B=256):
    g = gen()
    levels = []
    for s in (13, 26, 52):
        x = torch.randn((B, 3, 25, s, s), generator=g, device=DEV)
        x[:, :, 2:4] *= 0.5
        x[:, :, 4] = x[:, :, 4] * 2.5 - 10.0
        x[:, :, 5:] = x[:, :, 5:] * 2.0 - 3.0
        levels.append(x.reshape(B, 75, s, s))
    ls = ops.make_levels(levels)
    anchors = YOLOV3_LEVEL_ANCHORS
    c = ops.yolov3_decode_filter(ls, 20, anchors, (416, 416), 0.001, max_cand=16384)
    ms_dec = timed(lambda: ops.yolov3_decode_filter(ls, 20, anchors, (416, 416), 0.001, max_cand=16384), iters)

    def full():
        cc = ops.yolov3_decode_filter(ls, 20, anchors, (416, 416), 0.001, max_cand=16384)
        ops.sort_nms(cc, 0.5, ops.RULE_PER_CLASS, ops.ORDER_CLASS_MAJOR, max_det=0, max_out=8192)
    ms_tot = timed(full, iters)
    return report("yolov3_voc", B, 10647 * 25 * 4, ms_dec, ms_tot, {"cand_per_image": float(c.count.float().mean())})


def bench_head_fused(iters, B=64):
    """SURVEY §8f rank 3 at the C2 shape: the head's last 1x1 convolutions fused with decode + filter (tcgen05) vs the
    unfused pair (cuDNN convolutions writing the (B,144,A) head + the streaming decode kernel re-reading it)."""
    g = gen()
    sizes = ((80, 80), (40, 40), (20, 20))
    bf = [torch.randn((B, 64, h, w), generator=g, device=DEV) for h, w in sizes]
    cf = [torch.randn((B, 80, h, w), generator=g, device=DEV) for h, w in sizes]
    bw = [torch.randn((64, 64), generator=g, device=DEV) * 0.375 for _ in sizes]          # box logits ~ N(0, 3^2)
    cw = [torch.randn((80, 80), generator=g, device=DEV) * 0.4825 for _ in sizes]         # class logits ~ N(-18.19, 4.3155^2)
    bb = [torch.zeros((64,), device=DEV) for _ in sizes]
    cb = [torch.full((80,), -18.19, device=DEV) for _ in sizes]
    strides = (8.0, 16.0, 32.0)
    fused = lambda: ops.yolov8_head_decode_filter(bf, cf, bw, bb, cw, cb, strides, 0.001)   # noqa: E731
    c = fused()
    ms_eager = timed(fused, iters)
    ms_fused = timed_graph(fused, iters)
    w4 = [(a[:, :, None, None].contiguous(), b[:, :, None, None].contiguous()) for a, b in zip(bw, cw)]

    def conv_only():
        return [torch.cat((torch.nn.functional.conv2d(x, wa, ba), torch.nn.functional.conv2d(y, wb, bb_)), 1)
                for x, y, (wa, wb), ba, bb_ in zip(bf, cf, w4, bb, cb)]

    def unfused():
        return ops.yolov8_decode_filter(ops.make_levels(conv_only(), strides), 80, 0.001)
    ms_conv = timed(conv_only, iters)
    ms_unfused = timed(unfused, iters)
    cand = float(c.count.float().mean())
    return report("yolov8_head_fused_C2", B, 144 * 8400 * 4 + 24 * cand, ms_fused, ms_fused,
                  {"cand_per_image": cand, "eager_call_ms": round(ms_eager, 4), "unfused_ms": round(ms_unfused, 4), "unfused_conv_cat_ms": round(ms_conv, 4),
                   "speedup_vs_unfused": round(ms_unfused / ms_fused, 2),
                   "note": "inputs = the features BEFORE the last 1x1 convs; bytes = (c2 + c3) x A x 4 per image; the unfused "
                           "path additionally writes and re-reads the 4.84 MB/image head (torch conv2d = cuDNN, allow_tf32 default)"})


TABLE = {"head_fused": bench_head_fused, "yolov8": bench_yolov8, "centernet": bench_centernet, "ssd": bench_ssd, "yolov7": bench_yolov7,
         "yolov3": bench_yolov3}


def collect(iters=30, only=("centernet", "ssd", "yolov7", "yolov3", "head_fused"), device=None):
    """bench.py's `paths` key: one dict per configuration, measured in the calling process."""
    global DEV
    if device is not None:
        DEV = torch.device(device)
    out = []
    for name in only:
        out.append(TABLE[name](iters))
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--only", default="yolov8,centernet,ssd,yolov7,yolov3")
    a = ap.parse_args()
    for name in a.only.split(","):
        print(json.dumps(TABLE[name](a.iters)), flush=True)
