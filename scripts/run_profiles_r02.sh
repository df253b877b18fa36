# profiles of round 2 (run on the GPU box, every ncu command only after the same command exited 0 without ncu)
set -x
out=gpurun_out
timeout 400 python scripts/kernel_zoo.py --time > $out/r02_zoo_time.log 2>&1; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $out/r02_zoo.csv python scripts/kernel_zoo.py > $out/r02_zoo_ncu.log 2>&1; echo rc=$?
# launch list of the bench command (both workloads of the default run, shortened year)
timeout 300 python bench.py --nsteps 24 --steps 1 --warmup 1 --no-cpu-baseline > $out/r02_bench24.log 2>&1; echo rc=$?
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r02_launches_bench_nsteps24.csv python bench.py --nsteps 24 --steps 1 --warmup 1 --no-cpu-baseline > $out/r02_ncu_ll.log 2>&1; echo rc=$?
# full captures of the two step kernels
timeout 300 python bench.py --nsteps 12 --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $out/r02_plain_f.log 2>&1; echo rc=$?
timeout 500 ncu --set full --clock-control none --import-source on -k regex:step_fused_kernel -c 1 -o $out/r02_prof_fused -f python bench.py --nsteps 12 --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $out/r02_ncu_f.log 2>&1; echo rc=$?
timeout 300 python bench.py --module phosphorus --nsteps 24 --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $out/r02_plain_p3.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_fused_p3 -c 1 -o $out/r02_prof_p3 -f python bench.py --module phosphorus --nsteps 24 --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $out/r02_ncu_p3.log 2>&1; echo rc=$?
# full captures of the round-2 kernels: fused Gram-Schmidt, lin_comb, panel substitution
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mgs|lin_comb|banded_solve_panel" -c 8 -o $out/r02_prof_new -f python scripts/kernel_zoo.py > $out/r02_ncu_new.log 2>&1; echo rc=$?
tail -30 $out/r02_zoo_time.log
