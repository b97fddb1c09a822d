mkdir -p gpurun_out
L=gpurun_out/r2_h6.log
: > $L
timeout 600 python -m pytest tests/test_head_fused_gpu.py -q -m gpu --timeout=300 2>&1 | tail -5 >> $L
for sp in 2 0; do echo "split $sp" >> $L; CVPP_HEAD_SPLIT=$sp timeout 200 python tools/bench_paths.py --only head_fused --iters 50 2>&1 | tail -3 | cut -c1-130 >> $L; done
for dbg in 2 3; do echo "dbg $dbg" >> $L; CVPP_HEAD_DEBUG=$dbg timeout 200 python tools/bench_paths.py --only head_fused --iters 50 2>&1 | tail -3 | cut -c1-130 >> $L; done
cat $L
CVPP_HEAD_DEBUG=0 timeout 200 python tools/head_fused_timing.py > gpurun_out/r2_head_timing_dbg0.log 2>&1; grep "decode_ms\|kernel cycles\|lag\|landed (as\|steady\|^end\|^first" gpurun_out/r2_head_timing_dbg0.log; sed -n 50,58p gpurun_out/r2_head_timing_dbg0.log
