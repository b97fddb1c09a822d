timeout 900 python -m pytest tests/test_yolov3_gpu.py tests/test_yolov7_gpu.py tests/test_eval_gpu.py -q -m gpu --timeout=300 2>&1 | tail -3
timeout 200 python tools/bench_paths.py --only yolov7,yolov3 --iters 40 2>&1 | cut -c1-200
