CVPP_BENCH_BS1_EARLY=1 timeout 600 python bench.py --steps 50 --warmup 5 --no-paths --no-c5 --no-reference-gpu --no-cpu > gpurun_out/r2_w2_bench.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_w2_bench.json").read().strip().splitlines()[-1])
print(json.dumps(d["bs1_latency"]))
PY
