"""bs=1 latency breakdown (BASELINE configs[0] shape, conf .25): graph replays of decode only, NMS only, both; L2 flushed before
every replay, CUDA events around the replay."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from computervision.pytorch_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1234)
levels = []
for h, w in ((80, 80), (40, 40), (20, 20)):
    x = torch.randn((1, 144, h, w), generator=g, device=dev)
    x[:, :64] *= 3.0; x[:, 64:] *= 4.3155; x[:, 64:] += -18.19
    levels.append(x)
ls = ops.make_levels(levels, (8.0, 16.0, 32.0))
post = ops.Yolov8Postprocessor(1, 8400, 80, dev)
flush = torch.empty((192 * 1024 * 1024,), dtype=torch.uint8, device=dev)
def graph_of(fn):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    return gr
def p50(gr, do_flush=True, n=60):
    lat = []
    for i in range(n):
        if do_flush: flush.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
        if i >= 10: lat.append(a.elapsed_time(b) * 1e3)
    lat.sort(); return lat[len(lat) // 2]
c = ops.yolov8_decode_filter(ls, 80, 0.25)
empty = graph_of(lambda: None) if False else None
g_dec = graph_of(lambda: ops.yolov8_decode_filter(ls, 80, 0.25))
g_nms = graph_of(lambda: ops.sort_nms(c, 0.7, max_det=300, max_nms=30000))
g_all = graph_of(lambda: post(ls, 0.25, 0.7))
noop = torch.zeros((1,), device=dev)
g_one = graph_of(lambda: noop.add_(1))
for name, gr in (("one tiny kernel", g_one), ("decode+filter (memset + kernel)", g_dec), ("sort+NMS", g_nms), ("postprocess (memset, decode, NMS)", g_all)):
    print(f"{name:40s} L2 flushed {p50(gr):6.2f} us   warm {p50(gr, False):6.2f} us")

# the bench's exact bs = 1 input: the first image of its 64-image synthetic batch
import bench as _b
lv64 = _b.synth_levels_device(torch, dev, 1234, 64)
ls1 = ops.make_levels([l[:1].contiguous() for l in lv64], (8.0, 16.0, 32.0))
post1 = ops.Yolov8Postprocessor(1, 8400, 80, dev, max_det=300)
g1 = post1.capture(ls1, 0.25, 0.7)
print(f"{'bench input, GraphedPostprocess':40s} L2 flushed {p50(g1.graph):6.2f} us   warm {p50(g1.graph, False):6.2f} us   kept {int(post1.det.count.item())} cands {int(post1.det.cand_count.item())}")
print("cands of the tool's own input:", int(c.count.item()))
