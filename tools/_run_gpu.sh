timeout 200 python tools/nms_phase_timing_v7.py 2>&1 | tail -3
