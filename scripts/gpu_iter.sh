#!/bin/bash
# One GPU iteration on the label kernel: the label parity tests, a short device-resident bench and
# (optionally) an ncu --set full capture.  usage: gpu_iter.sh TAG [ncu]
TAG=$1
python -m pytest tests -m gpu -x -q -k "label or hypothesis or full_size or index_reads" 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e > gpurun_out/b_$TAG.json 2> gpurun_out/b_$TAG.err || tail -5 gpurun_out/b_$TAG.err
python -c "
import json; d=json.load(open('gpurun_out/b_$TAG.json')); print('value %.2f G/s  step %.2f ms  kernel %.2f ms' % (d['value']/1e9, d['ms_per_step'], d['roofline']['kernel_ms']))"
if [ "$2" = "ncu" ]; then
  ncu --set full --clock-control none --import-source on -k regex:labelseq -s 3 -c 1 -o gpurun_out/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_$TAG.log 2>&1
fi
