mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_yolov3_gpu.py tests/test_eval_gpu.py -q -m gpu --timeout=300 2>&1 | tail -3
for w in 24 20; do CVPP_YA_WARPS=$w timeout 200 python tools/bench_paths.py --only yolov3 --iters 50 2>&1 | cut -c1-130; done
