#!/bin/bash
# run on the GPU box: bench every build/variants/*.so (device-resident arm only)
cp spacegraphcats_b200/libsgc_b200.so /tmp/base.so
for f in build/variants/*.so; do
  v=$(basename $f .so)
  cp $f spacegraphcats_b200/libsgc_b200.so
  python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --no-query > gpurun_out/bv_$v.json 2> gpurun_out/bv_$v.err || tail -3 gpurun_out/bv_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bv_$v.json')); print('$v: value %.2f G/s  step %.2f ms  kernel %.2f ms' % (d['value']/1e9, d['ms_per_step'], d['roofline']['kernel_ms']))"
done
cp /tmp/base.so spacegraphcats_b200/libsgc_b200.so
