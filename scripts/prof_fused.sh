#!/bin/bash
# one ncu --set full capture of the persistent fused step kernel (48 time steps in one launch) per variant
# usage: scripts/prof_fused.sh <tag> [bench args...]   (environment selects the variant)
tag=$1; shift
cmd="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --nsteps 48 $*"
$cmd > gpurun_out/plain_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_fused -c 1 -o gpurun_out/prof_$tag -f $cmd > gpurun_out/ncu_$tag.log 2>&1
echo "prof $tag rc=$?"
