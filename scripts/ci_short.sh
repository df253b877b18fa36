#!/bin/bash
# port of the reference's scripts/ci_short.sh (its lint / pytest / cfg-usage steps belong to the reference's
# own source tree and are not repeated): test_problem iage + phosphorus, one fixed-point iteration
source "$(dirname "$0")/ci_common.sh"

echo running setup_solver
$cli setup_solver --fp_cnt 1 --depth_nlevs 20 --persist --model_name test_problem \
    --workdir $HOME/ci_short_workdir --deprecation_warning_to_error "$@" || err_cnt=$((err_cnt+1))

baseline_cmp $HOME/ci_short_workdir $baselines/ci_short depth_axis.nc
for fname in fcn_00.nc hist_00.nc init_iterate.nc init_iterate_00.nc; do
    baseline_cmp $HOME/ci_short_workdir/gen_init_iterate $baselines/ci_short $fname
done

echo err_cnt=$err_cnt
exit $err_cnt
