mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_epilogue_gpu.py tests/test_yolov8_gpu.py tests/test_sort_nms_gpu.py -q -m gpu --timeout=300 2>&1 | tail -5
for mode in fused eager graph; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --no-c5 --gather-mode $mode > gpurun_out/r2_l1_bench_2gpu_$mode.json 2> gpurun_out/r2_l1_bench_2gpu_$mode.err; tail -2 gpurun_out/r2_l1_bench_2gpu_$mode.err | grep -v "^\*\|OMP"; python - <<PY
import json
d=json.loads(open("gpurun_out/r2_l1_bench_2gpu_$mode.json").read().strip().splitlines()[-1])
print("$mode", "value", d["value"], "ms", d["ms_per_step"], "barrier-each", d.get("ms_per_step_barrier_each_step"), "verified", d.get("gather_verified"), "host", d["host_issue_ms_per_step"], "launches", d["gpu_launches"])
PY
done
