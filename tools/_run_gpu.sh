# round 2, call 1: new tests first (fail-fast off so that every failure is seen), smoke, bench, probes
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -q -m gpu -x --timeout=600 > gpurun_out/r2_c1_tests.log 2>&1; tail -15 gpurun_out/r2_c1_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_c1_smoke.log 2>&1; tail -1 gpurun_out/r2_c1_smoke.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_c1_bench.json 2> gpurun_out/r2_c1_bench.err; tail -5 gpurun_out/r2_c1_bench.err
cut -c1-400 gpurun_out/r2_c1_bench.json
timeout 200 python tools/h2d_probe.py > gpurun_out/r2_c1_h2d.log 2>&1; tail -2 gpurun_out/r2_c1_h2d.log
timeout 200 python tools/h2d_probe.py --numa > gpurun_out/r2_c1_h2d_numa.log 2>&1; tail -2 gpurun_out/r2_c1_h2d_numa.log
nvidia-smi topo -m > gpurun_out/r2_c1_topo.log 2>&1; lscpu | head -30 >> gpurun_out/r2_c1_topo.log
