"""Runs the fused sort+NMS kernel a few times on the bench workload's candidates (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from computervision.pytorch_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1234)
levels = []
for h, w in ((80, 80), (40, 40), (20, 20)):
    x = torch.randn((64, 144, h, w), generator=g, device=dev)
    x[:, :64] *= 3.0; x[:, 64:] *= 4.3155; x[:, 64:] += -18.19
    levels.append(x)
ls = ops.make_levels(levels, (8.0, 16.0, 32.0))
c = ops.yolov8_decode_filter(ls, 80, 0.001)
for _ in range(4):
    det = ops.sort_nms(c, 0.7, max_det=300, max_nms=30000)
torch.cuda.synchronize()
print("kept", det.count.float().mean().item())
