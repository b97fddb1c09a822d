run() { echo "== $*"; env "$@" timeout 200 python tools/bench_paths.py --only yolov8 --iters 100 2>&1 | cut -c1-120; }
run A=1
run CVPP_DECODE_CPL=4
run CVPP_DECODE_STAGES=3 CVPP_DECODE_WARPS=12
run CVPP_DECODE_STAGES=3 CVPP_DECODE_WARPS=10
run CVPP_DECODE_STAGES=2 CVPP_DECODE_WARPS=14
run CVPP_DECODE_STAGES=2 CVPP_DECODE_WARPS=16
