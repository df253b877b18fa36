"""CPU test of the derived hist variables (nk_ooc_b200/hist.py): fed with the tracer
snapshots of the reference's committed hist files, the derived variables reproduce the ones the
reference wrote next to them (py_driver_2d/tracer_module_state.py:214-260,
test_problem/tracer_module_state.py:166-199)."""
import os

import numpy as np
import pytest

from nk_ooc_b200 import hist
from nk_ooc_b200.spatial_axis import SpatialAxis


@pytest.fixture(scope="module")
def base(golden_dir):
    return np.load(os.path.join(golden_dir, "baselines.npz"))


def test_py_driver_2d_derived_hist_variables(base):
    pre = "ci_py_driver_2d_iage/"
    depth = SpatialAxis("depth", base[pre + "grid_vars/depth_edges"])
    ypos = SpatialAxis("ypos", base[pre + "grid_vars/ypos_edges"])
    vals = base[pre + "hist_0000/iage"]
    assert vals.shape == (61, 30, 30)
    got = hist.derived_values("iage", vals, depth, ypos)
    for suff in ("time_mean", "time_std", "time_delta", "depth_int", "ypos_mean", "depth_ypos_int"):
        want = base[pre + f"hist_0000/iage_{suff}"]
        np.testing.assert_allclose(got[f"iage_{suff}"], want, rtol=1e-13, atol=1e-13 * np.abs(want).max(), err_msg=suff)
    np.testing.assert_allclose(got["iage_time_anom"], vals - got["iage_time_mean"], rtol=0, atol=0)
    names = [s[0] for s in hist.derived_specs("iage", {"long_name": "ideal age", "units": "years"}, depth, ypos)]
    assert names == ["iage_time_mean", "iage_time_anom", "iage_time_std", "iage_time_delta", "iage_depth_int",
                     "iage_ypos_mean", "iage_depth_ypos_int"]


def test_test_problem_derived_hist_variables(base):
    pre = "ci_short/"
    depth = SpatialAxis("depth", base[pre + "depth_axis/depth_edges"])
    for name in ("iage", "po4", "po4_uptake", "po4_s_restore_tau_r"):
        vals = base[pre + f"hist_00/{name}"]
        assert vals.shape == (101, 20)
        got = hist.derived_values(name, vals, depth)
        for suff in ("time_mean", "time_std", "time_delta", "depth_int"):
            want = base[pre + f"hist_00/{name}_{suff}"]
            np.testing.assert_allclose(got[f"{name}_{suff}"], want, rtol=1e-13, atol=1e-13 * np.abs(want).max(),
                                       err_msg=f"{name}_{suff}")
    w = hist.time_mean_weights(101)
    assert abs(w.sum() - 1.0) < 1e-15 and w[0] == 0.5 * w[1]
