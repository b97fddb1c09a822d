"""Debug probe of the fused head kernel: dump the materialised head and compare with a float64 contraction."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import test_head_fused_gpu as T
from computervision.pytorch_b200 import ops
np.set_printoptions(linewidth=200)
B, sizes, c2, c3, nc = 2, ((16, 16),), 64, 80, 80
bf, cf, bw, cw, bb, cb = T._dyadic_case(3, B, sizes, c2, c3, nc, -14.0)
head = T._materialised_head(bf, cf, bw, cw, bb, cb)[0].reshape(B, 144, -1)
cand, got = ops.yolov8_head_decode_filter(T._to(bf), T._to(cf), T._to(bw), T._to(bb), T._to(cw), T._to(cb), (8.0,), 0.001, return_head=True)
torch.cuda.synchronize()
got = got.cpu().numpy()
print("count", cand.count.tolist())
err = np.abs(got - head)
print("max err", err.max(), "box part", err[:, :64].max(), "cls part", err[:, 64:].max())
print("want[0,:4,:6]\n", head[0, :4, :6], "\ngot\n", got[0, :4, :6])
print("want cls[0,64:68,:6]\n", head[0, 64:68, :6], "\ngot\n", got[0, 64:68, :6])
ok = err < 1e-6
print("fraction exact: box", ok[:, :64].mean(), "cls", ok[:, 64:].mean())
print("exact by cell (first 40):", ok[0, :64].mean(0)[:40].round(2))
print("exact by channel:", ok[0].mean(1).round(2))
print("bias-only? ", np.abs(got[0, :64] - bb[0][:, None]).max())
# partial-K hypotheses: only the first k channels contributed?
for kk in (8, 16, 32, 64):
    part = np.einsum("bkhw,nk->bnhw", bf[0][:, :kk].astype(np.float64), bw[0][:, :kk].astype(np.float64)).reshape(B, 64, -1) + bb[0][None, :, None]
    print("first", kk, "channels only:", np.abs(got[:, :64] - part).max())
