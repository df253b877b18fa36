# profiles of round 1 (run on the GPU box): per-kernel DRAM table, launch list of the bench command, full capture of
# the phosphorus step kernel
set -x
timeout 300 python scripts/kernel_zoo.py --time > gpurun_out/zoo_time.log 2>&1; echo rc=$?
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/zoo.csv python scripts/kernel_zoo.py > gpurun_out/zoo_ncu.log 2>&1; echo rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_fused.csv python bench.py --nsteps 24 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_ll2.log 2>&1; echo rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_p3.csv python bench.py --module phosphorus --nsteps 24 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_ll3.log 2>&1; echo rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_fused_p3 -c 1 -o gpurun_out/prof_p3c -f python bench.py --module phosphorus --nsteps 48 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_p3c.log 2>&1; echo rc=$?
cat gpurun_out/zoo_time.log
