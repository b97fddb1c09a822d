timeout 900 python -m pytest tests/test_ssd_gpu.py tests/test_eval_gpu.py -q -m gpu --timeout=300 2>&1 | tail -3
timeout 200 python tools/bench_paths.py --only ssd --iters 50 2>&1 | cut -c1-230
