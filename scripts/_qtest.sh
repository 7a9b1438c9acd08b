# GPU-box helper: GPU parity tests + a short device-resident bench (label and query arms); prints value / step / kernel ms
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --no-cpu --steps 5 --no-e2e 2>gpurun_out/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value']/1e9, d['ms_per_step'], d['roofline']['kernel_ms'])
q=d['query']; print(q.get('value',0)/1e9, q.get('ms_per_step'), q.get('e2e',{}).get('value'), q.get('hit_fraction'), q.get('distinct_fraction'), q.get('pairs_per_query'), q.get('error'))"
tail -3 gpurun_out/err.log
