"""Per-phase clock64 breakdown of the fused sort+NMS kernel (needs libcvpp_timing.so built with -DCVPP_NMS_TIMING)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from computervision.pytorch_b200 import _lib
_lib.SO_PATH = os.path.join(ROOT, "computervision", "pytorch_b200", "libcvpp_timing.so")
import numpy as np, torch
from computervision.pytorch_b200 import ops
dev = torch.device("cuda:0")
BS = int(sys.argv[1]) if len(sys.argv) > 1 else 64
CONF = float(sys.argv[2]) if len(sys.argv) > 2 else 0.001
g = torch.Generator(device=dev); g.manual_seed(1234)
levels = []
for h, w in ((80, 80), (40, 40), (20, 20)):
    x = torch.randn((BS, 144, h, w), generator=g, device=dev)
    x[:, :64] *= 3.0; x[:, 64:] *= 4.3155; x[:, 64:] += -18.19
    levels.append(x)
ls = ops.make_levels(levels, (8.0, 16.0, 32.0))
c = ops.yolov8_decode_filter(ls, 80, CONF)
for _ in range(3):
    det = ops.sort_nms(c, 0.7, max_det=300, max_nms=30000)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 1024)()
_lib.lib().cvpp_debug_n2_timing(buf)
t16 = np.array(buf[:]).reshape(64, 16)[:BS]
t = t16[:, :8]
print('  class sorts', ((t16[:,8]-t16[:,2])/1.965e3).mean(), 'big classes', ((t16[:,9]-t16[:,8])/1.965e3).mean(), 'gather', ((t16[:,3]-t16[:,9])/1.965e3).mean())
print('  prefix hist+bin', ((t16[:,10]-t16[:,0])/1.965e3).mean(), 'class hist+segments', ((t16[:,1]-t16[:,10])/1.965e3).mean())
print('  suppress_all', ((t16[:,11]-t16[:,3])/1.965e3).mean(), 'survivor count', ((t16[:,4]-t16[:,11])/1.965e3).mean())
d = np.diff(t, axis=1) / 1.965e3   # us at 1965 MHz
names = ["hist+scan", "scatter", "class sort+gather", "suppress", "select", "final sort", "output"]
for i, nme in enumerate(names):
    print(f"{nme:20s} mean {d[:, i].mean():7.2f} us   max {d[:, i].max():7.2f} us")
print("total", (t[:, 7] - t[:, 0]).mean() / 1.965e3, "us; cands", c.count.float().mean().item())
