set -x
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 1500 gpurun_out/bench_default.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_fused.csv python bench.py --nsteps 24 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_ll2.log 2>&1
tail -c 300 gpurun_out/ncu_ll2.log
