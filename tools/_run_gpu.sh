timeout 1500 python -m pytest tests -q -m gpu --timeout=300 2>&1 | tail -3
timeout 400 python tools/bench_paths.py --only ssd,yolov7,yolov3,head_fused,centernet --iters 40 2>&1 | cut -c1-150
