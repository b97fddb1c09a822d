set -x
python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_r1_s5.log 2>&1; tail -2 gpurun_out/gpu_tests_r1_s5.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1_s5.log 2>&1; tail -1 gpurun_out/smoke_r1_s5.log
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r1_s5.json 2> gpurun_out/bench_r1_s5.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r1_s5.json 2>&1
cat gpurun_out/bench_r1_s5.json | cut -c1-250; cat gpurun_out/bench_ref_r1_s5.json | cut -c1-200
