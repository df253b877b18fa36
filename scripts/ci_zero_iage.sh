#!/bin/bash
# port of the reference's scripts/ci_zero_iage.sh: Newton-Krylov from an all-zero iterate
source "$(dirname "$0")/ci_common.sh"
workdir=$HOME/ci_zero_iage_workdir
opts="--model_name test_problem --depth_nlevs 20 --tracer_module_names iage --workdir $workdir"

echo running setup_solver for zero iage
$cli setup_solver --fp_cnt 0 --persist --init_iterate_opt zeros $opts --deprecation_warning_to_error "$@" \
    || err_cnt=$((err_cnt+1))

echo running nk_driver for zero iage
$cli nk_driver $opts "$@" || err_cnt=$((err_cnt+1))

echo err_cnt=$err_cnt
exit $err_cnt
