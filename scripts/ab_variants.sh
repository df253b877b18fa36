#!/bin/bash
# A/B builds of libnkb200.so for kernel experiments: every argument "tag:-DFLAG1 -DFLAG2" gives
# newton-krylov_ooc_b200/variants/libnkb200_<tag>.so (git-ignored, travels with gpurun); "base:" is the plain build.
# usage: scripts/ab_variants.sh "base:" "nosleep:-DP3_PSLEEP_NS=0"
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
csrc=$root/newton-krylov_ooc_b200/csrc
out=$root/newton-krylov_ooc_b200/variants
mkdir -p $out
for spec in "$@"; do
  tag=${spec%%:*}; flags=${spec#*:}
  tmp=$(mktemp -d)
  for f in nkb_api nkb_tables nkb_stage nkb_stage_tma nkb_step_fused nkb_column nkb_ops nkb_krylov nkb_banded; do
    if [ "$f" = nkb_step_fused ] || [ "$f" = nkb_banded ]; then
      /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I$root/include $flags -c $csrc/$f.cu -o $tmp/$f.o &
    else
      cp $csrc/$f.o $tmp/$f.o
    fi
  done
  wait
  /usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out/libnkb200_$tag.so $tmp/*.o -lcudart
  rm -rf $tmp
  echo "built $out/libnkb200_$tag.so ($flags)"
done
