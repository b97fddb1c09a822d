"""Timeline of one YOLOv8 post-processing step (kineto/CUPTI): kernel start/duration and the gaps between
them, to see what the CUDA-event step time is made of.  Usage: python tools/trace_step.py [--graph]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from computervision.pytorch_b200 import ops
import bench

dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1234)
levels = []
for h, w in bench.SIZES:
    x = torch.randn((bench.BS, 144, h, w), generator=g, device=dev)
    x[:, :64] *= 3.0; x[:, 64:] *= 4.3155; x[:, 64:] += -18.19
    levels.append(x)
ls = ops.make_levels(levels, bench.STRIDES)
post = ops.Yolov8Postprocessor(bench.BS, bench.A, 80, dev)
use_graph = "--graph" in sys.argv
gp = post.capture(ls, bench.CONF, bench.IOU) if use_graph else None
step = (lambda: gp.replay()) if use_graph else (lambda: post(ls, bench.CONF, bench.IOU))
for _ in range(10): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
prev_end = None
for e in evs:
    st, en = e.time_range.start - t0, e.time_range.end - t0
    gap = (st - prev_end) if prev_end is not None else 0
    print(f"{st:10.1f}us  dur {en-st:8.1f}us  gap {gap:7.1f}us  {e.name[:70]}")
    prev_end = en
