set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_head_fused_gpu.py -q -m gpu --timeout=120 > gpurun_out/r2_c4_tests.log 2>&1; tail -30 gpurun_out/r2_c4_tests.log
timeout 200 python tools/bench_paths.py --only head_fused,yolov8 --iters 50 > gpurun_out/r2_c4_head_paths.jsonl 2>&1; cat gpurun_out/r2_c4_head_paths.jsonl
