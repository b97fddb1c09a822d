mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_yolov3_gpu.py -q -m gpu --timeout=300 2>&1 | tail -3
for w in 24 22 20; do CVPP_YA_WARPS=$w timeout 200 python tools/bench_paths.py --only yolov3 --iters 50 2>&1 | cut -c1-130; done
timeout 300 ncu --set full --clock-control none --import-source on -f -k "regex:yolo_anchor_stream" --launch-skip 7 --launch-count 1 -o gpurun_out/r2_n_v3 python tools/bench_paths.py --only yolov3 --iters 1 > gpurun_out/r2_n_v3.log 2>&1; tail -1 gpurun_out/r2_n_v3.log | cut -c1-100
