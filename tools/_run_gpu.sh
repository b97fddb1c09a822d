timeout 900 python -m pytest tests/test_sort_nms_gpu.py tests/test_nms_golden_gpu.py tests/test_properties_gpu.py tests/test_yolov7_gpu.py tests/test_ssd_gpu.py tests/test_yolov8_gpu.py -q -m gpu --timeout=300 2>&1 | tail -2
timeout 300 python tools/bench_paths.py --only yolov7,yolov3,ssd,yolov8 --iters 30 2>&1 | cut -c1-110
