timeout 600 python -m pytest tests/test_epilogue_gpu.py tests/test_yolov8_gpu.py tests/test_yolov7_gpu.py -q -m gpu --timeout=300 2>&1 | tail -8
