timeout 200 python tools/bench_paths.py --only head_fused --iters 50 2>&1 | cut -c1-130
for ns in 160 400; do cp computervision/pytorch_b200/libcvpp_ns$ns.so computervision/pytorch_b200/libcvpp.so; echo "sleep $ns"; timeout 200 python tools/bench_paths.py --only head_fused --iters 50 2>&1 | cut -c1-130; done
