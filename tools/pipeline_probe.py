"""Throughput probe: K steps of decode+NMS replayed from CUDA graphs on ONE stream vs alternating over TWO streams
(two buffer sets: the NMS of batch k overlaps the decode of batch k+1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from bench_paths import gen, DEV
from computervision.pytorch_b200 import ops

g = gen(); B = 64; K = 200
levels = []
for h, w in ((80, 80), (40, 40), (20, 20)):
    x = torch.randn((B, 144, h, w), generator=g, device=DEV)
    x[:, :64] *= 3.0; x[:, 64:] *= 4.3155; x[:, 64:] += -18.19
    levels.append(x)
ls = ops.make_levels(levels, (8.0, 16.0, 32.0))
posts = [ops.Yolov8Postprocessor(B, 8400, 80, DEV) for _ in range(2)]
graphs = [pp.capture(ls, 0.001, 0.7) for pp in posts]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def run(two):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for s in streams:
        s.wait_event(e0)
    for k in range(K):
        i = k & 1 if two else 0
        with torch.cuda.stream(streams[i]):
            graphs[i].replay()
    for s in streams:
        main.wait_stream(s)
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K * 1e3


for _ in range(2):
    run(True); run(False)
print(f"dynamic={os.environ.get('CVPP_DECODE_DYNAMIC', '1')}: one stream {run(False):.1f} us/step, two streams {run(True):.1f} us/step")
ref = posts[0].det.count.clone(); torch.cuda.synchronize()
print("kept", posts[0].det.count.float().mean().item(), posts[1].det.count.float().mean().item(),
      "identical sets:", bool(torch.equal(posts[0].det.anchor, posts[1].det.anchor)))
