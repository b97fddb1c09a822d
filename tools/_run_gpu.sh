timeout 1500 python -m pytest tests -q -m gpu -x --timeout=300 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 --no-paths --no-c5 --no-reference-gpu --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',round(d['value']),'frac',round(d['roofline']['frac'],3),'bs1',d['bs1_latency']['p50_us'],round(d['bs1_latency']['mean_us'],2),d['bs1_latency']['min_us'],'e2e',round(d['e2e']['value']))"
