#!/bin/bash
# port of the reference's scripts/ci_py_driver_2d_iage_column_regions.sh: 20 x 3 grid without lateral processes,
# every ypos column its own region, Newton-Krylov to convergence
source "$(dirname "$0")/ci_common.sh"
workdir=$HOME/ci_py_driver_2d_iage_column_regions_workdir
base=$baselines/ci_py_driver_2d_iage_column_regions
opts="--model_name py_driver_2d --tracer_module_names iage --depth_nlevs 20 --ypos_nlevs 3 --max_abs_vvel 0.0 --horiz_mix_coeff 0.0 --workdir $workdir"

echo running setup_solver
$cli setup_solver --fp_cnt 1 --persist $opts --deprecation_warning_to_error "$@" || err_cnt=$((err_cnt+1))

baseline_cmp $workdir $base grid_vars.nc
# hist file: its *_time_anom variables (x - time mean of x) are compared with the tolerance of x itself, see
# nk_ooc_b200/utils.py:isclose_all_vars — the one deviation from the reference's script
for fname in fcn_0000.nc hist_0000.nc init_iterate.nc init_iterate_0000.nc; do
    baseline_cmp $workdir/gen_init_iterate $base $fname --atol 1.0e-6 --rtol 1.0e-3 --anom_suffix _time_anom
done

echo running nk_driver for py_driver_2d
$cli nk_driver $opts "$@" || err_cnt=$((err_cnt+1))

baseline_cmp $workdir/krylov_00 $base precond_00.nc
baseline_cmp $workdir/krylov_00 $base precond_fcn_00.nc --rtol 2.0e-3
baseline_cmp $workdir/krylov_00 $base basis_00.nc --atol 5.0e-5
baseline_cmp $workdir/krylov_00 $base perturb_fcn_w_raw_00.nc --atol 5.0e-6
baseline_cmp $workdir/krylov_00 $base krylov_res_00.nc --rtol 1.9e-2
for fname in increment_00.nc iterate_01.nc; do
    baseline_cmp $workdir $base $fname --rtol 1.9e-2
done
newton_state_cmp $workdir $base

echo err_cnt=$err_cnt
exit $err_cnt
