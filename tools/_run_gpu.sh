set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1_s5.log 2>&1; tail -1 gpurun_out/smoke_r1_s5.log
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r1_s5.json 2> gpurun_out/bench_r1_s5.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r1_s5.json 2>&1
python tools/bench_paths.py --iters 50 > gpurun_out/paths_r1_s5.jsonl 2> gpurun_out/paths_r1_s5.err
python bench.py --steps 3 --warmup 3 --no-cpu --no-graph > gpurun_out/plain_s5.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_s5.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-graph > gpurun_out/ncu_s5a.log 2>&1
cat gpurun_out/bench_r1_s5.json | cut -c1-300; cat gpurun_out/bench_ref_r1_s5.json | cut -c1-200
