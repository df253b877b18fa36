"""CPU tests of the solver persistence (Newton_state.json / *_stats.nc), no GPU needed"""
import json

import numpy as np
import pytest
from scipy.io import netcdf_file

from nk_ooc_b200 import solver_state as ss


def test_step_log_strings_and_resume_rewind(tmp_path):
    """same strings as the reference's state files (baselines/ci_long_iage/Newton_state.json):
    per-iteration steps are prefixed NN:, inc_iteration is logged under the new iteration"""
    st = ss.SolverState("Newton", str(tmp_path))
    st.log_step("Newton iterate 0 written", per_iteration=False)
    st.log_step("comp_fcn complete for X/fcn_00.nc")
    st.log_step("comp_fcn complete for X/fcn_00.nc")  # idempotent
    assert st.inc_iteration() == 1
    st.log_step("comp_fcn complete for X/fcn_01.nc")
    saved = json.load(open(tmp_path / "Newton_state.json"))
    assert saved == {"iteration": 1, "step_log": ["__init__", "Newton iterate 0 written",
                                                  "00:comp_fcn complete for X/fcn_00.nc", "01:inc_iteration",
                                                  "01:comp_fcn complete for X/fcn_01.nc"]}
    back = ss.SolverState("Newton", str(tmp_path), resume=True)
    assert back.get_iteration() == 1 and back.step_logged("comp_fcn complete for X/fcn_01.nc")
    assert not back.step_logged("comp_fcn complete for X/fcn_00.nc")  # other iteration
    rew = ss.SolverState("Newton", str(tmp_path), resume=True, rewind=True)
    assert not rew.step_logged("comp_fcn complete for X/fcn_01.nc")
    assert rew.step_was_rewound("comp_fcn complete for X/fcn_01.nc")
    with pytest.raises(RuntimeError):
        ss.SolverState("Newton", str(tmp_path), resume=False, rewind=True)


def test_saved_values_roundtrip_exactly(tmp_path):
    st = ss.SolverState("Krylov", str(tmp_path))
    h_mat = np.random.default_rng(0).normal(size=(2, 3, 2, 4))
    st.set_value_saved_state("h_mat", h_mat)
    st.set_value_saved_state("armijo_ind", 3)
    back = ss.SolverState("Krylov", str(tmp_path), resume=True)
    np.testing.assert_array_equal(back.get_value_saved_state("h_mat"), h_mat)
    assert back.get_value_saved_state("armijo_ind") == 3
    raw = json.load(open(tmp_path / "Krylov_state.json"))
    assert set(raw["h_mat"]) == {"__ndarray__"}  # the reference's tagging (solver_state.py:149-166)


class _State:
    def __init__(self, mean, norm):
        self._m, self._n = np.asarray(mean), np.asarray(norm)

    def mean(self):
        return self._m

    def norm(self):
        return self._n


def test_stats_file_layout_growth_and_resume(tmp_path):
    mods = [("iage", "years"), ("phosphorus", None)]
    sf = ss.StatsFile("Newton", str(tmp_path), 3, mods, ss.NEWTON_VARS)
    it0 = _State(np.arange(6.0).reshape(2, 3), 10 + np.arange(6.0).reshape(2, 3))
    sf.put(0, iterate=it0, fcn=it0)
    sf.put(0, increment=it0, Krylov_iterations=4, increment_scalef=np.ones((2, 3)), Armijo_factor=0.5 * np.ones((2, 3)))
    sf.put(1, iterate=it0)
    with netcdf_file(str(tmp_path / "Newton_stats.nc"), "r", mmap=False) as f:
        assert f.version_byte == 2
        assert f.dimensions["iteration"] is None and f.dimensions["region"] == 3
        names = set(f.variables)
        for key in ("iterate", "fcn", "increment"):
            for method in ("mean", "norm"):
                for mod, _ in mods:
                    assert f"{key}_{method}_{mod}" in names
        assert {"increment_scalef_iage", "Armijo_factor_phosphorus", "Krylov_iterations", "iteration", "region"} <= names
        np.testing.assert_array_equal(f.variables["iteration"].data, [0, 1])
        np.testing.assert_array_equal(f.variables["region"].data, [0, 1, 2])
        np.testing.assert_array_equal(f.variables["iterate_norm_phosphorus"].data, [[13, 14, 15], [13, 14, 15]])
        assert f.variables["iterate_mean_iage"].units == b"years"
        assert f.variables["iterate_mean_iage"].long_name == b"mean of iage Newton iterate"
        assert not hasattr(f.variables["iterate_mean_phosphorus"], "units")
        # iteration 1 of the variables not yet written holds the fill value (stats_file.py:129-139)
        assert (f.variables["fcn_mean_iage"].data[1] == ss.FILL_F8).all()
        assert f.variables["Krylov_iterations"].data[1] == ss.FILL_I4
        np.testing.assert_array_equal(f.variables["Krylov_iterations"].data[:1], [4])
    again = ss.StatsFile("Newton", str(tmp_path), 3, mods, ss.NEWTON_VARS, resume=True)
    again.put(1, fcn=it0)
    with netcdf_file(str(tmp_path / "Newton_stats.nc"), "r", mmap=False) as f:
        np.testing.assert_array_equal(f.variables["fcn_mean_iage"].data, [[0, 1, 2], [0, 1, 2]])
        np.testing.assert_array_equal(f.variables["Armijo_factor_iage"].data[0], [0.5, 0.5, 0.5])
    kf = ss.StatsFile("Krylov", str(tmp_path), 3, mods, ss.KRYLOV_VARS)
    kf.put_invariant(precond_rhs_norm=np.ones((2, 3)))
    kf.put(0, precond_resid_norm=2 * np.ones((2, 3)))
    with pytest.raises(RuntimeError):
        kf.put_invariant(precond_resid_norm=np.ones((2, 3)))
    with netcdf_file(str(tmp_path / "Krylov_stats.nc"), "r", mmap=False) as f:
        assert f.variables["precond_rhs_norm_iage"].dimensions == ("region",)
        assert f.variables["precond_resid_norm_iage"].dimensions == ("iteration", "region")


def test_hist_statistics_in_the_stats_file(tmp_path):
    """time mean with down-weighted end points and ypos mean of the tracer-like hist variables
    (py_driver_2d/tracer_module_state.py:281-345), coordinate variables copied from the hist file,
    fill values for iterations that have no hist statistics, reload on resume"""
    rng = np.random.default_rng(1)
    nt, nz, ny = 5, 4, 3
    vals = rng.normal(size=(nt, nz, ny))
    hist = str(tmp_path / "hist_00.nc")
    with netcdf_file(hist, "w", version=2) as f:
        f.createDimension("time", None)
        f.createDimension("depth", nz)
        f.createDimension("ypos", ny)
        t = f.createVariable("time", "f8", ("time",))
        d = f.createVariable("depth", "f8", ("depth",))
        d.units = "m"
        y = f.createVariable("ypos", "f8", ("ypos",))
        v = f.createVariable("iage", "f8", ("time", "depth", "ypos"))
        v.long_name = "ideal age"
        v.units = "years"
        v.cell_methods = "time: point"
        d[:] = np.arange(nz) + 0.5
        y[:] = np.arange(ny) * 2.0
        t[:nt] = np.arange(nt)
        v[:nt] = vals
    sf = ss.StatsFile("Newton", str(tmp_path), 1, [("iage", "years")], ss.NEWTON_VARS)
    widths = np.array([1.0, 2.0, 1.0])
    sf.put_hist_stats(1, hist, ["iage", "not_in_hist"], {"ypos": widths})
    w = np.array([0.5, 1, 1, 1, 0.5]) / 4.0
    want = np.einsum("i,ijk", w, vals)
    with netcdf_file(str(tmp_path / "Newton_stats.nc"), "r", mmap=False) as f:
        assert f.variables["iage"].dimensions == ("iteration", "depth", "ypos")
        assert f.variables["iage_mean_ypos"].dimensions == ("iteration", "depth")
        np.testing.assert_allclose(f.variables["iage"].data[1], want, rtol=1e-15)
        np.testing.assert_allclose(f.variables["iage_mean_ypos"].data[1], want @ (widths / 4.0), rtol=1e-15)
        assert (f.variables["iage"].data[0] == ss.FILL_F8).all()
        assert f.variables["iage"].units == b"years" and not hasattr(f.variables["iage"], "cell_methods")
        np.testing.assert_array_equal(f.variables["depth"].data, np.arange(nz) + 0.5)
        assert f.variables["depth"].units == b"m"
    again = ss.StatsFile("Newton", str(tmp_path), 1, [("iage", "years")], ss.NEWTON_VARS, resume=True)
    again.put_hist_stats(2, hist, ["iage"], {"ypos": widths})
    with netcdf_file(str(tmp_path / "Newton_stats.nc"), "r", mmap=False) as f:
        np.testing.assert_allclose(f.variables["iage"].data[1], want, rtol=1e-15)
        np.testing.assert_allclose(f.variables["iage"].data[2], want, rtol=1e-15)
        assert f.variables["iteration"].shape[0] == 3


class _RefStatsFile:
    """records the calls of the reference's StatsFile API (nk_ooc/stats_file.py:70-125)"""

    def __init__(self):
        self.dimensions, self.vars_metadata, self.invariant, self.per_iteration = {}, {}, {}, {}

    def def_dimensions(self, dimensions):
        self.dimensions.update(dimensions)

    def def_vars(self, vars_metadata, caller=None):
        self.vars_metadata.update(vars_metadata)

    def put_vars_iteration_invariant(self, name_vals_dict):
        self.invariant.update(name_vals_dict)

    def put_vars(self, iteration, name_vals_dict):
        self.per_iteration[iteration] = dict(name_vals_dict)


def test_stats_hooks_write_through_the_references_stats_file_api(tmp_path):
    """ModelStateBase.def_stats_vars / put_stats_vars* hand whatever stats file the solver owns to
    solver_state.as_hist_stats: under the reference's NewtonSolver that is the reference's StatsFile
    (newton_solver.py:52-58,330), served by RefStatsFileAdapter with the same variables and values as this
    package's own StatsFile"""
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from baseline_files import materialise

    hist = os.path.join(materialise(str(tmp_path / "baselines")), "ci_py_driver_2d_iage", "hist_0000.nc")
    with netcdf_file(hist, "r", mmap=False) as f:
        ypos_delta = np.array(f.variables["ypos_delta"].data)
        iage = np.array(f.variables["iage"].data)
    names, weights = ["iage", "iage_slow_rest"], {"ypos": ypos_delta}
    ref = _RefStatsFile()
    ad = ss.as_hist_stats(ref)
    assert isinstance(ad, ss.RefStatsFileAdapter)
    ad.def_hist_stats(hist, names, weights)
    ad.put_hist_coordinates(hist)
    ad.put_hist_stats(2, hist, names, weights)
    assert ref.dimensions == {"depth": 30, "ypos": 30}
    assert set(ref.vars_metadata) == {"depth", "ypos", "iage", "iage_mean_ypos", "iage_slow_rest",
                                      "iage_slow_rest_mean_ypos"}
    meta = ref.vars_metadata["iage_mean_ypos"]
    assert meta["dimensions"] == ("iteration", "depth") and meta["attrs"] == {"long_name": "ideal age", "units": "years"}
    assert set(ref.invariant) >= {"depth", "ypos"} and ref.invariant["depth"].shape == (30,)
    w = np.full(61, 1.0 / 60.0)
    w[[0, -1]] *= 0.5
    want = np.einsum("i,i...", w, iage)
    np.testing.assert_allclose(ref.per_iteration[2]["iage"], want, rtol=1e-14)
    np.testing.assert_allclose(ref.per_iteration[2]["iage_mean_ypos"], want @ (ypos_delta / ypos_delta.sum()), rtol=1e-13)
    # the same numbers as this package's own StatsFile
    own = ss.StatsFile("Newton", str(tmp_path), 1, [("iage", "years")], ss.NEWTON_VARS)
    assert ss.as_hist_stats(own) is own
    own.put_hist_stats(0, hist, names, weights)
    with netcdf_file(str(tmp_path / "Newton_stats.nc"), "r", mmap=False) as f:
        np.testing.assert_allclose(np.array(f.variables["iage_mean_ypos"].data)[0], ref.per_iteration[2]["iage_mean_ypos"],
                                   rtol=1e-14)
