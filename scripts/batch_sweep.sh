#!/bin/bash
# "report the smallest B/G at which efficiency drops" — members per GPU against throughput on ONE GPU
# (refined125x150, forced and phosphorus, 240 uniform steps per year; evals/s scale with 240/2640 to the graded schedule)
for module in forced phosphorus; do
  for m in 4096 2048 1024 512 256 128 64 32 16; do
    python bench.py --module $module --members $m --nsteps 240 --steps 4 --warmup 2 --no-cpu-baseline --no-extra 2>&1 | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$module', $m, round(d['value'],1), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"
  done
done
