"""Timeline of CTA 0 of the fused head kernel (instrumented build libcvpp_timing.so: make -C csrc timing)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from computervision.pytorch_b200 import _lib
_lib.SO_PATH = os.path.join(ROOT, "computervision", "pytorch_b200", "libcvpp_timing.so")
import numpy as np, torch
import bench_paths as bp
r = bp.bench_head_fused(20)
print({k: r[k] for k in ("decode_ms", "decode_frac_of_hbm_peak")})
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (8 * 512))()
_lib.lib().cvpp_debug_hf_timing(buf)
t = np.array(buf[:], dtype=np.int64).reshape(8, 512)
t0, t1 = t[5, 0], t[5, 1]
print("kernel cycles", t1 - t0)
n_chunks = int((t[2] > 0).sum()); n_tiles = int((t[3] > 0).sum())
print("chunks", n_chunks, "tiles", n_tiles)
rel = lambda a: a - t0
print("chunk:  tma_issued  operands_landed  committed   (landed - issued)  (commit - landed)   d(commit)")
for c in list(range(0, 12)) + list(range(50, 75)):
    if c < n_chunks:
        print(f"{c:4d} {rel(t[0,c]):10d} {rel(t[1,c]):10d} {rel(t[2,c]):10d}   {t[1,c]-t[0,c]:8d} {t[2,c]-t[1,c]:8d} {t[2,c]-t[2,c-1] if c else 0:8d}")
st = int(os.environ.get("CVPP_HEAD_STAGES", "8"))
lag = [t[0, c] - t[2, c - st] for c in range(st + 40, min(n_chunks, st + 100))]
print("stages", st, "commit(c - stages) -> reissue(c) lag: median", int(np.median(lag)), "min", min(lag), "max", max(lag))
lat = [t[1, c] - t[0, c] for c in range(40, min(n_chunks, 140))]
print("TMA issue -> landed (as seen by the MMA warp): median", int(np.median(lat)), "min", min(lat), "max", max(lat))
per = (t[2, 140] - t[2, 40]) / 100.0
print("steady-state cycles per chunk", per, "per tile", 5 * per)
print("tile: acc_full_seen  box_released (d)  class_released (d)  class_done (d)   d(seen)")
for i in range(0, min(n_tiles, 29)):
    print(f"{i:4d} {rel(t[3,i]):10d} {rel(t[4,i]):10d} {t[4,i]-t[3,i]:6d} {rel(t[6,i]):10d} {t[6,i]-t[3,i]:6d} {rel(t[7,i]):10d} {t[7,i]-t[3,i]:6d} {t[3,i]-t[3,i-1] if i else 0:8d}")

cb = (ctypes.c_ulonglong * (256 * 4))()
_lib.lib().cvpp_debug_hf_cta(cb)
c = np.array(cb[:], dtype=np.uint64).reshape(256, 4)[:148].astype(np.int64)
base = c[:, 0].min()
c -= base
print("per-CTA globaltimer (ns from the earliest CTA start): start / first operands / last accumulator committed / end")
print("start  min %d max %d" % (c[:, 0].min(), c[:, 0].max()))
print("first  min %d max %d" % (c[:, 1].min(), c[:, 1].max()))
print("lastacc min %d max %d" % (c[:, 2].min(), c[:, 2].max()))
print("end    min %d max %d median %d" % (c[:, 3].min(), c[:, 3].max(), np.median(c[:, 3])))
print("CTA 0:", c[0])
order = np.argsort(c[:, 3])
print("earliest finishers", order[:8], c[order[:8], 3])
print("latest finishers", order[-8:], c[order[-8:], 3])
