# end-of-round validation on the GPU box: full GPU test suite, the two bench workloads, strong-scaling member count,
# ncu capture of the phosphorus step kernel
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo rc=$?
timeout 300 python bench.py --module phosphorus --no-cpu-baseline > gpurun_out/bench_phosphorus.json 2> gpurun_out/bench_phosphorus.err; echo rc=$?
timeout 200 python bench.py --members 512 --no-cpu-baseline > gpurun_out/bench_b512.json 2> gpurun_out/bench_b512.err; echo rc=$?
timeout 200 python bench.py --module phosphorus --members 512 --no-cpu-baseline > gpurun_out/bench_p3_b512.json 2> gpurun_out/bench_p3_b512.err; echo rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_fused_p3 -c 1 -o gpurun_out/prof_p3d -f python bench.py --module phosphorus --nsteps 48 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_p3d.log 2>&1; echo rc=$?
python -c "
import json
for f in ('bench_default','bench_phosphorus','bench_b512','bench_p3_b512'):
    d=json.loads(open('gpurun_out/'+f+'.json').readline())
    print(f, round(d['value'],1), round(d['ms_per_step'],1), round(d['roofline']['frac'],3), round(d['e2e']['value'],1), d['clocks'])
"
