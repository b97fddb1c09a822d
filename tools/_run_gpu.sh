mkdir -p gpurun_out
for cfg in "a:" "b:--eager-gather" "c:--eager-gather --no-multicast" "d:"; do
tag=${cfg%%:*}; fl=${cfg#*:}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 50 --warmup 5 --no-c5 $fl > gpurun_out/r2_k2_bench_8gpu_$tag.json 2> gpurun_out/r2_k2_bench_8gpu_$tag.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r2_k2_bench_8gpu_$tag.json").read().strip().splitlines()[-1])
print("$tag $fl", "value", d["value"], "ms", d["ms_per_step"], "barrier-each", d.get("ms_per_step_barrier_each_step"), "verified", d.get("gather_verified"), "host", d["host_issue_ms_per_step"])
PY
done
timeout 300 python bench.py --steps 50 --warmup 5 --no-c5 --no-paths --no-reference-gpu --no-cpu > gpurun_out/r2_k2_bench_1gpu.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/r2_k2_bench_1gpu.json").read().strip().splitlines()[-1])
print("1 gpu same box", "value", d["value"], "ms", d["ms_per_step"])
PY
