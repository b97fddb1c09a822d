mkdir -p gpurun_out
for n in 8 4 2 1; do
if [ $n = 1 ]; then
timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 --no-paths --no-reference-gpu > gpurun_out/r2_z1_bench_${n}gpu.json 2> gpurun_out/r2_z1_bench_${n}gpu.err
else
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/r2_z1_bench_${n}gpu.json 2> gpurun_out/r2_z1_bench_${n}gpu.err
fi
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_z1_bench_${n}gpu.json").read().strip().splitlines()[-1])
print($n, "value", round(d["value"]), "us/step", round(1e3*d["ms_per_step"],2), "barrier-each", d.get("ms_per_step_barrier_each_step"), "verified", d.get("gather_verified"), "e2e", round(d["e2e"]["value"]), "ceiling", round(d["e2e"]["ceiling"]["value"]), "h2d/gpu", round(d["e2e"]["h2d_GBps_per_gpu"],1))
c=d["c5"]; print("   c5", round(c["images_per_s"]), c["ms_per_step"], c["compute_ms_per_step"], c["gather_verified"])
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29539 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 2>/dev/null | cut -c1-160
