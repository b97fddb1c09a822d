import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
from computervision.pytorch_b200 import ops
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda:0')
g = torch.Generator(device=dev); g.manual_seed(99)
B=64
pred = torch.empty((B,128,128,84), device=dev)
pred[..., :80] = torch.randn((B,128,128,80), generator=g, device=dev)*1.5-5.0
pred[..., 80:82] = torch.rand((B,128,128,2), generator=g, device=dev)
pred[..., 82:] = torch.rand((B,128,128,2), generator=g, device=dev)*20
for _ in range(3): ops.centernet_decode(pred, 100, 0.001)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ops.centernet_decode(pred, 100, 0.001); torch.cuda.synchronize()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and 'centernet' in e.name: print(f"{e.name[:50]:50s} {e.device_time:8.1f} us")
