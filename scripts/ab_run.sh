#!/bin/bash
# runs bench.py once per variant library (scripts/ab_variants.sh) on the same box, interleaved REPS times
# usage: scripts/ab_run.sh "<bench args>" tag1 tag2 ...      (REPS=2 by default)
args=$1; shift
reps=${REPS:-2}
for r in $(seq $reps); do
  for tag in "$@"; do
    NKB_LIB_PATH=$PWD/newton-krylov_ooc_b200/variants/libnkb200_$tag.so python bench.py $args --no-cpu-baseline --no-extra 2>&1 | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tag', round(d['value'],1), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done
