#!/bin/bash
# A/B of run-time switches on one box: scripts/env_ab.sh "<bench args>" "VAR=a" "VAR=b" ...   (REPS repetitions, interleaved)
args=$1; shift
for r in $(seq ${REPS:-2}); do
  for e in "$@"; do
    env $e python bench.py $args --no-cpu-baseline --no-extra 2>&1 | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$e', round(d['value'],1), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"
  done
done
