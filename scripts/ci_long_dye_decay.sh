#!/bin/bash
# port of the reference's scripts/ci_long_dye_decay.sh: two parameterised dye_decay modules
source "$(dirname "$0")/ci_common.sh"
workdir=$HOME/ci_long_dye_decay_workdir
opts="--model_name test_problem --depth_nlevs 20 --tracer_module_names dye_decay_{suff}:001:010 --newton_rel_tol 1.0e-6 --workdir $workdir"

echo running setup_solver for dye_decay
$cli setup_solver --fp_cnt 1 --persist $opts --deprecation_warning_to_error "$@" || err_cnt=$((err_cnt+1))

echo running nk_driver for dye_decay
$cli nk_driver $opts "$@" || err_cnt=$((err_cnt+1))

newton_state_cmp $workdir $baselines/ci_long_dye_decay

echo err_cnt=$err_cnt
exit $err_cnt
