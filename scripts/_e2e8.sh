# 8-GPU-box helper (profiles/r02_h2d_probe_8gpu.txt): the end-to-end arm with two copy streams, one copy stream, and without the label download
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 12 --warmup 3 --no-query --no-partitioned --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value']/1e9, d['e2e']['value']/1e9, d['e2e']['ms_per_step'])"; }
run default
SGC_STREAM_SERIAL=1 run serial
SGC_STREAM_NO_D2H=1 run no_d2h
