set -x
NKB_YG=1 python -m pytest tests/test_gpu_stage.py -x -q 2>&1 | tail -2
python -m pytest tests/test_gpu_stage.py -x -q 2>&1 | tail -2
run() { echo "== $*"; env "$@" python bench.py $CFG --nsteps 120 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('evals/s(scaled to S=2400)', d['value']*120/2400, 'frac', d['roofline']['frac'], 'avg_launch_ms', d['roofline']['avg_launch_ms'])
    else: print(l.strip()[:200])
"; }
CFG=""
run NKB_X=0
run NKB_MPT=1
run NKB_YG=1
run NKB_YG=1 NKB_MPT=1 NKB_BX=32 NKB_JT=8
run NKB_YG=1 NKB_MPT=1 NKB_BX=32 NKB_JT=4
run NKB_YG=1 NKB_MPT=2 NKB_BX=16 NKB_JT=8
CFG="--grid default40x50 --module iage"
run NKB_X=0
run NKB_MPT=1
run NKB_YG=1
run NKB_YG=1 NKB_MPT=1 NKB_BX=32 NKB_JT=8
run NKB_YG=1 NKB_MPT=1 NKB_BX=32 NKB_JT=4
run NKB_YG=1 NKB_MPT=2 NKB_BX=16 NKB_JT=8
