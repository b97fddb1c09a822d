#!/usr/bin/env python
"""bench.py — headline benchmark: YOLOv8 eval-mode decode + NMS (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path (decode+filter -> fused per-class sort + class-aware NMS) over one
batch of synthetic head tensors per GPU: bs=64, 640x640 (8400 anchors), 80 classes, reg_max=16,
conf .001, IoU .7, max_det 300.  N>1 is launched by torchrun (one rank per GPU, NCCL); every rank owns
its own batch (weak scaling, images shard naturally) and each step ends with the single all-gather of
the detections that the evaluation would need.

`value`      images/s over all GPUs, inputs already resident in HBM (CUDA events, max over ranks).
`e2e`        same metric through the public call with HOST buffers: pinned-host -> device copy of the
             head tensors and device -> host read of the detections inside the timed region.
`roofline`   the decode+filter kernel alone: algorithmic bytes / CUDA-event time vs measured HBM peak.
`cpu_baseline` the CPU oracle (a plain-C port of the reference path, oracle/) on this box's cores.
`--impl reference` times that CPU port as the reference arm (the reference itself is Python that
             cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "YOLOv8 decode+NMS images/sec at 1/2/4/8 B200; bs=1 p50 µs; decode HBM GB/s"
UNIT = "images/s"
BS, NC, REG_MAX = 64, 80, 16
SIZES = ((80, 80), (40, 40), (20, 20))
STRIDES = (8.0, 16.0, 32.0)
A = sum(h * w for h, w in SIZES)
CONF, IOU, MAX_DET = 0.001, 0.7, 300
HEAD_BYTES_PER_IMAGE = (4 * REG_MAX + NC) * A * 4          # 4 838 400
CAND_BYTES = 8 + 16                                        # key + xyxy box written per candidate
CONFIG = {
    "workload": "yolov8_eval_decode_nms: bs=64/GPU, 640x640 (8400 anchors), nc=80, reg_max=16, conf=0.001, "
                "iou=0.7, max_det=300 (BASELINE.json configs[1])",
    "batch_per_gpu": BS, "anchors": A, "conf": CONF, "iou": IOU, "max_det": MAX_DET,
    "l2": "inputs (310 MB/GPU) exceed the 126 MB L2; no flush needed between iterations",
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# -------------------------------------------------------------------------------------------------
# clocks
# -------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(busy), "reasons": sorted(reasons)}


# -------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference path)
# -------------------------------------------------------------------------------------------------
def cpu_port_images_per_s(levels_np, budget_s: float, min_reps: int = 2):
    """Times oracle decode + NMS (all host threads) on `levels_np`; returns (images/s, reps, threads)."""
    import oracle
    oracle.set_threads(0)
    n_img = levels_np[0].shape[0]
    # warm-up
    y = oracle.yolov8_decode(levels_np, STRIDES, NC)
    oracle.yolov8_nms(y, CONF, IOU, MAX_DET, nc=NC)
    reps, t0 = 0, time.perf_counter()
    while True:
        y = oracle.yolov8_decode(levels_np, STRIDES, NC)
        oracle.yolov8_nms(y, CONF, IOU, MAX_DET, nc=NC)
        reps += 1
        el = time.perf_counter() - t0
        if reps >= min_reps and el >= budget_s:
            break
        if el > 4 * budget_s:
            break
    return n_img * reps / el, reps, oracle.max_threads()


def synth_levels_numpy(seed: int, B: int):
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for h, w in SIZES:
        x = rng.standard_normal((B, 4 * REG_MAX + NC, h, w), dtype=np.float32)
        x[:, :4 * REG_MAX] *= np.float32(3.0)
        x[:, 4 * REG_MAX:] *= np.float32(4.3155)
        x[:, 4 * REG_MAX:] += np.float32(-18.19)
        out.append(x)
    return out


def run_reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    levels = synth_levels_numpy(1234, BS)
    import oracle
    oracle.set_threads(0)
    # W warm-up steps, then exactly K timed steps; one step = the whole 64-image batch
    for _ in range(args.warmup):
        oracle.yolov8_nms(oracle.yolov8_decode(levels, STRIDES, NC), CONF, IOU, MAX_DET, nc=NC)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.yolov8_nms(oracle.yolov8_decode(levels, STRIDES, NC), CONF, IOU, MAX_DET, nc=NC)
    el = time.perf_counter() - t0
    v = BS * args.steps / el
    cores = oracle.max_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CONFIG,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {BS} images (full batch), oracle/ C port of the reference "
                                   f"path on {cores} host threads"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from computervision.pytorch_b200 import ops

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("for --gpus N>1 launch with torch.distributed.run (one rank per GPU)")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # synthetic head, generated on the device (SURVEY §8d distribution), seed = base + rank
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    levels = []
    for h, w in SIZES:
        x = torch.randn((BS, 4 * REG_MAX + NC, h, w), generator=g, device=dev, dtype=torch.float32)
        x[:, :4 * REG_MAX] *= 3.0
        x[:, 4 * REG_MAX:] *= 4.3155
        x[:, 4 * REG_MAX:] += -18.19
        levels.append(x)
    ls = ops.make_levels(levels, STRIDES)
    post = ops.Yolov8Postprocessor(BS, A, NC, dev, max_det=MAX_DET)

    from computervision.pytorch_b200 import distributed as cvd
    if world > 1:
        # the detection all-gather of step k runs on a side stream and overlaps the decode of step k+1:
        # two alternating payload buffers [rows (B,300,7) | counts (B)], ONE collective per step
        side = torch.cuda.Stream(device=dev)
        n_pay = BS * MAX_DET * 7 + BS
        pay = [torch.empty((n_pay,), dtype=torch.float32, device=dev) for _ in range(2)]
        gout = [torch.empty((world, n_pay), dtype=torch.float32, device=dev) for _ in range(2)]
        gdone = [torch.cuda.Event() for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        for e in gdone:
            e.record()
        step_no = [0]
        peer = None
        if not args.nccl_gather:
            try:   # fused epilogue + all-gather over NVLink peer stores (no NCCL kernel next to the decode)
                peer = cvd.PeerGather(BS, MAX_DET, 7, dev, depth=max(2, args.pipeline_depth))
            except Exception as e:
                print(f"[bench] symmetric-memory peer gather unavailable ({e}); using NCCL all_gather", file=sys.stderr)
                peer = None

    graphed = None
    # Throughput mode (default): --pipeline-depth (3) detection buffer sets, each with its own CUDA graph and stream,
    # used round-robin - batches are independent, so the latency-bound NMS kernel of step k (64 CTAs) overlaps the
    # HBM-bound decode of the next steps (measured: depth 1 78 us/step, 2 60 us, 3 54 us, 4 54 us).
    # --no-pipeline runs strictly one step after the other on one stream.
    pipelined = (world == 1 or peer is not None) and not args.no_pipeline
    pipe = None
    if world == 1 or peer is not None:
        pipe = ops.PipelinedPostprocess(BS, A, NC, dev, ls, CONF, IOU, max_det=MAX_DET, depth=args.pipeline_depth if pipelined else 1,
                                        graph=not args.no_graph)
    if not args.no_graph:
        try:
            if pipe is not None:
                pass
            else:
                # one graph per payload buffer: postprocess + row packing (cvpp_detection_epilogue)
                post(ls, CONF, IOU)
                torch.cuda.synchronize()
                graphed = []
                for i in range(2):
                    gph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gph):
                        ops.detection_epilogue(post(ls, CONF, IOU), ops.ROWS_FULL, out=pay[i], packed=True)
                    graphed.append(gph)
        except Exception as e:  # report, fall back to the eager C call
            print(f"[bench] CUDA graph capture failed ({e}); using eager launches", file=sys.stderr)
            graphed = None

    def step():
        if world == 1:
            return pipe.submit()
        # (box 4, score, cls, anchor) as 7 fp32 columns + counts, delivered to every rank
        if peer is not None:
            # decode + NMS of step k, then its fused epilogue + peer-store all-gather, on pipeline stream k % 2
            # (buffer set and gather slot k % 2): both overlap the decode of step k + 1 on the other stream
            det = pipe.submit()
            with torch.cuda.stream(pipe.streams[pipe.last_slot]):
                ops.detection_epilogue_allgather(det, ops.ROWS_FULL, peer.peer_ptrs(pipe.last_slot), peer.rank)
            return det                                 # (the cross-rank barrier comes once, after the K steps)
        i = step_no[0] & 1
        step_no[0] += 1
        main = torch.cuda.current_stream()
        main.wait_event(gdone[i])                      # the gather that last read this buffer is done
        if graphed is not None:
            graphed[i].replay()
        else:
            ops.detection_epilogue(post(ls, CONF, IOU), ops.ROWS_FULL, out=pay[i], packed=True)
        ready[i].record(main)
        with torch.cuda.stream(side):
            side.wait_event(ready[i])
            dist.all_gather_into_tensor(gout[i], pay[i])
            gdone[i].record(side)
        return post.det

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed_plain(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if pipe is not None:
            pipe.fork()                                         # the pipeline streams start after e0 ...
        h0 = time.perf_counter()
        for _ in range(steps):
            fn()
        host_ms[0] = 1e3 * (time.perf_counter() - h0) / steps   # host time to ISSUE a step (diagnostic)
        if pipe is not None:
            pipe.join()                                         # ... and e1 waits for every one of them
        elif world > 1:
            torch.cuda.current_stream().wait_event(gdone[0])   # the side-stream gathers belong to the timed region
            torch.cuda.current_stream().wait_event(gdone[1])
        if world > 1 and peer is not None:
            peer.barrier(0)     # inside the timed region: every rank's rows have landed in every buffer
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    W, K = max(args.warmup, 3), args.steps
    for _ in range(W):
        step()
    barrier()   # also pays the one-time communicator set-up outside the measurement
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    # keep the timed region long enough for nvidia-smi to see it: repeat the K-step measurement
    ms_runs = []
    t_end = time.perf_counter() + 1.5
    while True:
        ms_runs.append(timed(step, K))
        stop = time.perf_counter() > t_end or len(ms_runs) >= 50
        if world > 1:
            f = torch.tensor([int(stop)], device=dev)
            dist.broadcast(f, 0)
            stop = bool(f.item())
        if stop:
            break
    ms = statistics.median(ms_runs)
    host_ms_step = host_ms[0]
    det = step()
    if pipe is not None:
        pipe.join()
    torch.cuda.synchronize()
    # the same K steps strictly one after the other on one stream (no overlap between steps), for comparison
    serial_ms = None
    if pipelined and world == 1 and not args.no_graph:
        sg = post.capture(ls, CONF, IOU)
        serial_ms = statistics.median([timed_plain(sg.replay, K) for _ in range(10)]) / K
    cand_mean = float(det.cand_count.float().mean().item())
    kept_mean = float(det.count.float().mean().item())

    # ---- roofline of the dominant kernel (decode+filter): CUDA events around every stage of the
    #      three-call pipeline (same kernels as the fused call), accumulated over K in-situ iterations
    # Each stage is launched K times back to back between two events on the launching stream (the
    # current torch stream): the average launch duration without the dependency gap of a mixed sequence.
    # (The stage is replayed from a one-launch CUDA graph: an eager call of the Python wrapper costs more host
    # time than the 25 us NMS kernel runs, which would time the host, not the kernel.)
    def burst(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if args.no_graph:
            replay = fn
        else:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            replay = g.replay
            replay()
        barrier()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(K):
            replay()
        b_.record()
        barrier()
        return a.elapsed_time(b_) / K

    ms_dec = burst(lambda: ops.yolov8_decode_filter(ls, NC, CONF))
    cand_fixed = ops.yolov8_decode_filter(ls, NC, CONF)
    ms_nms = burst(lambda: ops.sort_nms(cand_fixed, IOU, max_det=MAX_DET, max_nms=30000))
    if world > 1:
        t = torch.tensor([ms_dec, ms_nms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dec, ms_nms = (float(v) for v in t.tolist())
    clocks = sampler.stop() if rank == 0 else None

    # ---- bs=1 latency (BASELINE configs[0] shape: conf .25, IoU .7), CUDA-graph replay, L2 flushed
    #      before every timed replay (the 4.8 MB head would otherwise sit in the 126 MB L2)
    bs1 = None
    if rank == 0:
        lv1 = [l[:1].contiguous() for l in levels]
        ls1 = ops.make_levels(lv1, STRIDES)
        post1 = ops.Yolov8Postprocessor(1, A, NC, dev, max_det=MAX_DET)
        g1 = post1.capture(ls1, 0.25, IOU)
        flush = torch.empty((192 * 1024 * 1024,), dtype=torch.uint8, device=dev)
        lat = []
        for i in range(60):
            flush.fill_(i & 0xFF)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g1.replay()
            b.record()
            torch.cuda.synchronize()
            if i >= 10:
                lat.append(a.elapsed_time(b) * 1e3)
        lat.sort()
        bs1 = {"p50_us": lat[len(lat) // 2], "p95_us": lat[int(len(lat) * 0.95)], "reps": len(lat),
               "config": "bs=1, conf=0.25, iou=0.7 (BASELINE.json configs[0] shape), graph replay, L2 flushed",
               "kept": int(post1.det.count.item())}
        del flush

    # ---- e2e: pinned host buffers -> device -> kernels -> host
    host_levels = [torch.empty(l.shape, dtype=l.dtype).pin_memory() for l in levels]
    for hl, l in zip(host_levels, levels):
        hl.copy_(l)
    dev_in = [torch.empty_like(l) for l in levels]
    ls_in = ops.make_levels(dev_in, STRIDES)
    h_box = torch.empty((BS, MAX_DET, 4), dtype=torch.float32).pin_memory()
    h_score = torch.empty((BS, MAX_DET), dtype=torch.float32).pin_memory()
    h_cls = torch.empty((BS, MAX_DET), dtype=torch.int32).pin_memory()
    h_anchor = torch.empty((BS, MAX_DET), dtype=torch.int32).pin_memory()
    h_count = torch.empty((BS,), dtype=torch.int32).pin_memory()
    h2d = sum(l.numel() * 4 for l in levels)
    d2h = sum(t.numel() * t.element_size() for t in (h_box, h_score, h_cls, h_anchor, h_count))

    def e2e_step():
        for d, h in zip(dev_in, host_levels):
            d.copy_(h, non_blocking=True)
        dt = post(ls_in, CONF, IOU)
        h_box.copy_(dt.box, non_blocking=True)
        h_score.copy_(dt.score, non_blocking=True)
        h_cls.copy_(dt.cls, non_blocking=True)
        h_anchor.copy_(dt.anchor, non_blocking=True)
        h_count.copy_(dt.count, non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the caller reads the detections

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        alg_bytes = BS * (HEAD_BYTES_PER_IMAGE + CAND_BYTES * cand_mean)
        achieved = alg_bytes / (ms_dec * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "decode_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        cpu = None
        if world == 1 and not args.no_cpu:
            v, reps, cores = cpu_port_images_per_s(synth_levels_numpy(1234, BS), budget_s=args.cpu_seconds)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{reps} reps x {BS} images of the same workload (oracle/ C port, {cores} host threads)"}
        line = {
            "metric": METRIC, "value": world * BS * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(CONFIG, candidates_per_image=cand_mean, kept_per_image=kept_mean,
                           all_gather=("none (1 GPU)" if world == 1 else
                                       "fused into the epilogue kernel: rows (B,300,7) fp32 + counts stored into every "
                                       "rank's buffer over NVLink peer memory every step (alternating slots), on the step's pipeline "
                                       "stream; one symmetric-memory barrier at the end of the timed region"
                                       if peer is not None else
                                       "one NCCL all-gather per step of [rows (B,300,7) fp32 | counts], on a side stream "
                                       "overlapping the next step's decode"),
                           launch="eager C call" if args.no_graph else "CUDA graph replay of cvpp_yolov8_postprocess",
                           pipeline=(f"{args.pipeline_depth} batches in flight: steps go round-robin over {args.pipeline_depth} streams / "
                                     "detection buffer sets, so the NMS kernel of step k overlaps the decode of the next steps "
                                     "(ops.PipelinedPostprocess); every step does the full decode+NMS, the timed region ends "
                                     "when every stream has drained"
                                     if pipelined else "none: one step after the other on one stream"),
                           timing=f"median of {len(ms_runs)} back-to-back {K}-step CUDA-event measurements"),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "yolov8_decode_stream_kernel<FULL=0, CPL=2, STAGES=2> (decode+filter)",
                         "ms_per_launch": ms_dec, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src},
            "stages_ms": {"decode_filter": ms_dec, "fused_sort_nms": ms_nms},
            "bs1_latency": bs1,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "e2e": {"value": world * BS * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / K},
            # per step: decode+filter and fused sort+NMS (+ the epilogue / peer-store gather kernel when N > 1)
            "gpu_launches": (2 if world == 1 else 3) * K,
            "host_issue_ms_per_step": host_ms_step,
            "serial_ms_per_step": serial_ms,   # one stream, no overlap between steps (decode + NMS back to back)
            "overlap": ("NMS of step k under the decode of step k+1" if pipelined else None) if world == 1 else
                       "NMS + epilogue/all-gather of step k under the decode of step k+1",
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="launch the three kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--pipeline-depth", type=int, default=3, help="batches in flight in throughput mode")
    ap.add_argument("--no-pipeline", action="store_true", help="one step after the other on one stream (no overlap of step k's NMS with step k+1's decode)")
    ap.add_argument("--nccl-gather", action="store_true", help="N>1: use NCCL all_gather instead of the fused peer-store epilogue")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
