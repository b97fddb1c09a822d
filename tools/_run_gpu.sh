mkdir -p gpurun_out
rm -f gpurun_out/r2_x1_errors.jsonl
CVPP_ERR_LOG=gpurun_out/r2_x1_errors.jsonl timeout 1500 python -m pytest tests -q -m gpu --timeout=300 > gpurun_out/r2_x1_gpu_tests.txt 2>&1; tail -2 gpurun_out/r2_x1_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_x1_smoke.log 2>&1; tail -1 gpurun_out/r2_x1_smoke.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_x1_bench.json 2> gpurun_out/r2_x1_bench.err; tail -2 gpurun_out/r2_x1_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_x1_bench_reference_arm.json 2>> gpurun_out/r2_x1_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_x1_bench.json").read().strip().splitlines()[-1])
print("value",d['value'],"ms",d['ms_per_step'],"serial",d['serial_ms_per_step'],"bs1",d['bs1_latency']['p50_us'],d['bs1_latency']['floor_us'],"frac",d['roofline']['frac'],"stages",d['stages_ms'],"e2e",d['e2e']['value'], d['e2e']['ceiling']['value'], "cpu", d['cpu_baseline']['value'])
for p in d['paths']: print(p['path'], p['decode_ms'], p['decode_frac_of_hbm_peak'], p['total_ms'])
print("c5", d['c5']['images_per_s'], d['c5']['ms_per_step'], d['c5']['decode_ms'], d['c5']['nms_ms'], d['c5']['decode_frac'])
print("refgpu", d['reference_gpu']['value'], "clocks", d['clocks'])
r=json.loads(open("gpurun_out/r2_x1_bench_reference_arm.json").read().strip().splitlines()[-1]); print("ref arm", r['value'])
PY
