# what `gpurun -- 'bash tools/_run_gpu.sh'` runs on the GPU box: the -m gpu suite, smoke(), and both bench arms
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout=300 > gpurun_out/gpu_tests.txt 2>&1; tail -2 gpurun_out/gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_arm.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_reference_arm.json
