# re-capture with the round's FINAL kernels (after the L2-promotion / output-ring / integer-limiter / idle-warp changes)
set -x
out=gpurun_out
timeout 300 python bench.py --nsteps 24 --steps 1 --warmup 1 --no-cpu-baseline > $out/r02f_bench24.log 2>&1; echo rc=$?
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r02f_launches_bench_nsteps24.csv python bench.py --nsteps 24 --steps 1 --warmup 1 --no-cpu-baseline > $out/r02f_ncu_ll.log 2>&1; echo rc=$?
timeout 300 python bench.py --nsteps 12 --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $out/r02f_plain_f.log 2>&1; echo rc=$?
timeout 500 ncu --set full --clock-control none --import-source on -k regex:step_fused_kernel -c 1 -o $out/r02f_prof_fused -f python bench.py --nsteps 12 --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $out/r02f_ncu_f.log 2>&1; echo rc=$?
timeout 300 python bench.py --module phosphorus --nsteps 24 --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $out/r02f_plain_p3.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_fused_p3 -c 1 -o $out/r02f_prof_p3 -f python bench.py --module phosphorus --nsteps 24 --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $out/r02f_ncu_p3.log 2>&1; echo rc=$?
