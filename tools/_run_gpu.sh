for w in 24 22; do CVPP_YA_WARPS=$w timeout 200 python tools/bench_paths.py --only yolov7 --iters 30 2>&1 | cut -c1-130; done
cp computervision/pytorch_b200/libcvpp_s4.so computervision/pytorch_b200/libcvpp.so
for w in 10 12 13; do echo "stages 4 warps $w"; CVPP_YA_WARPS=$w timeout 200 python tools/bench_paths.py --only yolov7 --iters 30 2>&1 | cut -c1-130; done
