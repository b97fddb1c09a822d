"""Throughput probe: K steps of decode+NMS through ops.PipelinedPostprocess at several pipeline depths
(depth 1 = one stream, strictly serial steps)."""
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
for depth in (1, 2, 3, 4):
    pipe = ops.PipelinedPostprocess(B, 8400, 80, DEV, ls, 0.001, 0.7, depth=depth)
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipe.fork()
        for k in range(K):
            pipe.submit()
        pipe.join()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / K * 1e3)
    print(f"depth {depth}: {best:.1f} us/step ({B / best * 1e3:.0f} K img/s)")
