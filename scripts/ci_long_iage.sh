#!/bin/bash
# port of the reference's scripts/ci_long_iage.sh: test_problem iage, Newton-Krylov to convergence
source "$(dirname "$0")/ci_common.sh"
workdir=$HOME/ci_long_iage_workdir
opts="--model_name test_problem --depth_nlevs 20 --tracer_module_names iage --workdir $workdir"

echo running setup_solver for iage
$cli setup_solver --fp_cnt 1 --persist $opts --deprecation_warning_to_error "$@" || err_cnt=$((err_cnt+1))

echo comparing iage from gen_init_iterate fixed point iteration to same from from ci_short
python - $HOME/ci_short_workdir/gen_init_iterate/hist_00.nc $workdir/gen_init_iterate/hist_00.nc <<'PY' || err_cnt=$((err_cnt+1))
import sys
import numpy as np
from scipy.io import netcdf_file
vals = []
for fname in sys.argv[1:3]:
    with netcdf_file(fname, "r", mmap=False) as fptr:
        vals.append(np.array(fptr.variables["iage"].data))
sys.exit(0 if np.array_equal(vals[0], vals[1]) else 1)
PY

echo running nk_driver for iage
$cli nk_driver $opts "$@" || err_cnt=$((err_cnt+1))

for fname in precond_00.nc precond_fcn_00.nc basis_00.nc perturb_fcn_w_raw_00.nc; do
    baseline_cmp $workdir/krylov_00 $baselines/ci_long_iage $fname
done
for fname in w_raw_00.nc w_00.nc krylov_res_00.nc; do
    baseline_cmp $workdir/krylov_00 $baselines/ci_long_iage $fname --rtol 2.0e-4
done
for fname in increment_00.nc iterate_01.nc; do
    baseline_cmp $workdir $baselines/ci_long_iage $fname --rtol 2.0e-4
done
newton_state_cmp $workdir $baselines/ci_long_iage

echo err_cnt=$err_cnt
exit $err_cnt
