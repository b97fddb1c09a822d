mkdir -p gpurun_out
L=gpurun_out/r2_h4.log
: > $L
timeout 900 python -m pytest tests/test_centernet_gpu.py tests/test_topk_gpu.py -q -m gpu --timeout=300 2>&1 | tail -5 >> $L
timeout 200 python tools/bench_paths.py --only centernet --iters 50 2>&1 | cut -c1-330 >> $L
cat $L
