#!/bin/bash
# port of the reference's scripts/ci_py_driver_2d_iage.sh: lat-depth iage on a 30 x 30 grid (the reference writes
# an override.cfg next to its input/py_driver_2d cfg files; here the same two values are command-line overrides
# of the built-in restatement of those cfg files), one fixed-point iteration
source "$(dirname "$0")/ci_common.sh"
workdir=$HOME/ci_py_driver_2d_iage_workdir

echo running setup_solver
$cli setup_solver --fp_cnt 1 --model_name py_driver_2d --tracer_module_names iage \
    --depth_nlevs 30 --ypos_nlevs 30 --workdir $workdir --deprecation_warning_to_error "$@" || err_cnt=$((err_cnt+1))

baseline_cmp $workdir $baselines/ci_py_driver_2d_iage grid_vars.nc
# hist file: its *_time_anom variables (x - time mean of x) are compared with the tolerance of x itself, see
# nk_ooc_b200/utils.py:isclose_all_vars — the one deviation from the reference's script
for fname in fcn_0000.nc hist_0000.nc init_iterate.nc init_iterate_0000.nc; do
    baseline_cmp $workdir/gen_init_iterate $baselines/ci_py_driver_2d_iage $fname --atol 1.0e-6 --rtol 1.0e-3 --anom_suffix _time_anom
done

echo err_cnt=$err_cnt
exit $err_cnt
