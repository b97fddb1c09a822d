cp computervision/pytorch_b200/libcvpp.so /tmp/new.so
for v in new prev new prev; do
if [ $v = prev ]; then cp computervision/pytorch_b200/libcvpp_prevnms.so computervision/pytorch_b200/libcvpp.so; else cp /tmp/new.so computervision/pytorch_b200/libcvpp.so; fi
timeout 300 python bench.py --steps 50 --warmup 5 --no-paths --no-c5 --no-reference-gpu --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v','us/step',round(1e3*d['ms_per_step'],2),'serial',round(1e3*d['serial_ms_per_step'],2),'nms',round(1e3*d['stages_ms']['fused_sort_nms'],2),'dec',round(1e3*d['stages_ms']['decode_filter'],2),'bs1',round(d['bs1_latency']['mean_us'],2))"
done
