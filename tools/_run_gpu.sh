timeout 600 python -m pytest tests/test_centernet_gpu.py tests/test_topk_gpu.py tests/test_eval_gpu.py -q -m gpu --timeout=300 2>&1 | tail -3
timeout 200 python tools/bench_paths.py --only centernet --iters 50 2>&1 | cut -c1-230
