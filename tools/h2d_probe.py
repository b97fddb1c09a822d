"""H2D bandwidth probe for the e2e leg: 310 MB (one C2 batch) from host to device - ordinary pinned memory vs
write-combined pinned memory (cudaHostAlloc), one copy vs split over 2/4 streams, NUMA-bound vs not."""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

if "--numa" in sys.argv:
    print("numa:", bench.bind_to_gpu_numa_node(0))
dev = torch.device("cuda:0")
N = 309_657_600
dst = torch.empty((N,), dtype=torch.uint8, device=dev)
rt = ctypes.CDLL("libcudart.so.12")


def host_alloc(flags):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(N), ctypes.c_uint(flags)) == 0
    buf = (ctypes.c_uint8 * N).from_address(p.value)
    t = torch.frombuffer(buf, dtype=torch.uint8)
    return t


def rate(src, streams, reps=8):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    chunk = (N + streams - 1) // streams
    best = 0.0
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i, s in enumerate(ss):
            with torch.cuda.stream(s):
                dst[i * chunk:(i + 1) * chunk].copy_(src[i * chunk:(i + 1) * chunk], non_blocking=True)
        torch.cuda.synchronize()
        best = max(best, N / (time.perf_counter() - t0) / 1e9)
    return round(best, 2)


out = {}
pinned = torch.empty((N,), dtype=torch.uint8).pin_memory()
pinned.fill_(3)
for s in (1, 2, 4):
    out[f"pinned_{s}stream"] = rate(pinned, s)
try:
    wc = host_alloc(4)   # cudaHostAllocWriteCombined
    wc.fill_(5)
    for s in (1, 2):
        out[f"writecombined_{s}stream"] = rate(wc, s)
except Exception as e:
    out["writecombined"] = f"failed: {e}"
print(json.dumps(out))
