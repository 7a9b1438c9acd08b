# GPU-box helper: everything behind profiles/r02_*: GPU tests, default bench, reference arm, ncu launch list, ncu --set full of the label kernel
set -x
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/f_pytest.log 2>&1; tail -3 gpurun_out/f_pytest.log
(time timeout 600 python bench.py) > gpurun_out/f_bench.log 2> gpurun_out/f_bench.err; tail -c 600 gpurun_out/f_bench.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/f_bench_ref.log 2> gpurun_out/f_bench_ref.err; tail -c 300 gpurun_out/f_bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --query-steps 2 > gpurun_out/f_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:labelseq -s 3 -c 1 -o gpurun_out/f_prof_label -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-query > gpurun_out/f_ncu_label.log 2>&1
ls -la gpurun_out/f_*
