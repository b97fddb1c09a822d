"""Per-phase clock64 breakdown of the fused sort+NMS kernel on the YOLOv7 C5-shard candidates (class-major, uncapped).
Needs libcvpp_timing.so (make -C computervision/pytorch_b200/csrc timing)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from computervision.pytorch_b200 import _lib
_lib.SO_PATH = os.path.join(ROOT, "computervision", "pytorch_b200", "libcvpp_timing.so")
import numpy as np, torch
import oracle
from computervision.pytorch_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(99)
B = 64
levels = []
for s in (20, 40, 80):
    x = torch.randn((B, 3, 85, s, s), generator=g, device=dev)
    x[:, :, 4] = x[:, :, 4] * 3.0 - 9.0
    x[:, :, 5:] = x[:, :, 5:] * 2.0 - 1.0
    levels.append(x.reshape(B, 255, s, s))
ls = ops.make_levels(levels)
cc = ops.yolov7_decode_filter(ls, 80, oracle.yolov7_level_anchors(), (640, 640), 0.001)
for _ in range(3):
    det = ops.sort_nms(cc, 0.3, ops.RULE_PER_CLASS, ops.ORDER_CLASS_MAJOR, max_det=0, max_out=8192)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 1024)()
_lib.lib().cvpp_debug_n2_timing(buf)
t = np.array(buf[:]).reshape(64, 16) / 1.965e3
def d(a, b): return float((t[:, b] - t[:, a]).mean())
print(f"cands/img {cc.count.float().mean().item():.0f}  kept {det.count.float().mean().item():.0f}")
print(f"class hist+segments {d(0,1):.1f}  scatter {d(1,2):.1f}  class sorts {d(2,8):.1f}  big classes {d(8,9):.1f}  gather {d(9,3):.1f}")
print(f"suppress_all {d(3,11):.1f}  (mark 11->4 {d(11,4):.1f})  output (4->7) {d(4,7):.1f}  total {d(0,7):.1f} us")
