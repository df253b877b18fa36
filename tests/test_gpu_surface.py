"""GPU tests of the operator-surface members beyond the hot-path operators: the per-module TracerModuleState
hooks with the reference's callback signatures (comp_tend / comp_jacobian, device backed), class discovery by
module path, log_vals, the stats-variable methods, reversed operators and the precond-matrix bookkeeping
(nk_ooc/model_state_base.py:113-180,201,310,379-402,627-667)."""
import logging
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _info2d(tmp, nz, ny, names, extra=None):
    info = {
        "model_name": "py_driver_2d", "tracer_module_names": names, "grid_vars_fname": os.path.join(tmp, "grid_vars.nc"),
        "depth_axisname": "depth", "depth_units": "m", "depth_edge_start": "0.0", "depth_edge_end": "4000.0",
        "depth_nlevs": str(nz), "depth_delta_ratio_max": "19.0",
        "ypos_axisname": "ypos", "ypos_units": "m", "ypos_edge_start": "0.0", "ypos_edge_end": "50.0e5",
        "ypos_nlevs": str(ny), "ypos_delta_ratio_max": "1.0", "max_abs_vvel": "0.1", "horiz_mix_coeff": "1000.0",
        "reinvoke": "False",
    }
    info.update(extra or {})
    return info


def test_py_driver_2d_tracer_module_hooks_match_reference(golden_dir, tmp_path):
    """iage / phosphorus: TracerModuleState.comp_tend(time, flat, processes) and comp_jacobian with the
    reference's callback signatures against the reference's own values (tests/golden/py_driver_2d.npz)"""
    from nk_ooc_b200.model_state_base import get_model_state_class
    from nk_ooc_b200.py_driver_2d.setup_solver import gen_grid_vars_file

    g = np.load(os.path.join(golden_dir, "py_driver_2d.npz"))
    tag, nz, ny = "g14x11", 14, 11
    ModelState = get_model_state_class("py_driver_2d")
    assert ModelState.__module__ == "nk_ooc_b200.py_driver_2d.model_state"
    info = _info2d(str(tmp_path), nz, ny, "iage,phosphorus")
    gen_grid_vars_file(info)
    ModelState.configure(info)
    try:
        ms = ModelState("zeros")
        names = [type(t).__name__ for t in ms.tracer_modules]
        assert names == ["iage", "phosphorus"]
        times = g[f"{tag}/times"]
        for tms, key in zip(ms.tracer_modules, ("iage", "phosphorus")):
            x = g[f"{tag}/{key}/x"]
            for i in range(4):
                got = tms.comp_tend(times[i], x.reshape(-1), ModelState.transport)
                assert isinstance(got, np.ndarray) and got.shape == (x.size,)
                want = g[f"{tag}/{key}/tend"][i].reshape(-1)
                np.testing.assert_allclose(got, want, rtol=0, atol=1e-12 * np.abs(want).max())
            got = tms.comp_jacobian(times[3], x.reshape(-1), ModelState.transport).toarray()
            want = g[f"{tag}/{key}/jac_dense_t3"]
            np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12 * np.abs(want).max())
        # bookkeeping of the preconditioner matrices (model_state_base.py:379-402)
        assert ms.precond_matrix_list() == ["phosphorus"]
        assert ms.hist_vars_for_precond_list() == ["po4", "time"]  # own entries, then those of "base"
        assert ms.tracer_names_per_precond_matrix() == {"phosphorus": ["po4"]}
    finally:
        ModelState.reset()


def test_test_problem_tracer_module_hooks_match_reference(golden_dir, tmp_path):
    from nk_ooc_b200.model_state_base import get_model_state_class
    from nk_ooc_b200.spatial_axis import spatial_axis_from_defn
    from nk_ooc_b200.test_problem.model_state import gen_depth_axis_file

    g = np.load(os.path.join(golden_dir, "test_problem.npz"))
    ModelState = get_model_state_class("test_problem")
    info = {"model_name": "test_problem", "tracer_module_names": "iage,dye_decay_{suff}:010,phosphorus",
            "po4_s_restoring_opt": "1", "grid_vars_fname": str(tmp_path / "depth_axis.nc"), "depth_axisname": "depth",
            "reinvoke": "False"}
    gen_depth_axis_file(info, spatial_axis_from_defn("depth", nlevs=20))
    ModelState.configure(info)
    try:
        ms = ModelState("zeros")
        assert [type(t).__name__ for t in ms.tracer_modules] == ["iage", "dye_decay", "phosphorus"]
        for tms in ms.tracer_modules:
            x = g[f"{tms.name}/x"]
            for i, t in enumerate(g["times"]):
                got = tms.comp_tend(t, x.reshape(-1), None)
                want = g[f"{tms.name}/tend"][i].reshape(-1)
                np.testing.assert_allclose(got, want, rtol=0, atol=1e-12 * np.abs(want).max(), err_msg=tms.name)
        assert ms.hist_vars_for_precond_list() == ["po4_s_restore_tau_r:mean", "mixing_coeff:mean", "mixing_coeff:log_mean"] \
            or set(ms.hist_vars_for_precond_list()) == {"po4_s_restore_tau_r:mean", "mixing_coeff:mean", "mixing_coeff:log_mean"}
        assert ms.tracer_modules[2].stats_vars_tracer_like()[-1] == "po4_uptake"
    finally:
        ModelState.reset()


def test_log_vals_reversed_operators_and_stats_methods(tmp_path, caplog):
    from scipy.io import netcdf_file

    from nk_ooc_b200 import solver_state
    from nk_ooc_b200.py_driver_2d.model_state import ModelState
    from nk_ooc_b200.py_driver_2d.setup_solver import gen_grid_vars_file

    nz, ny = 10, 3
    info = _info2d(str(tmp_path), nz, ny, "iage", {"max_abs_vvel": "0.0", "horiz_mix_coeff": "0.0"})
    gen_grid_vars_file(info)
    ModelState.configure(info, steps_per_year=120)
    try:
        rng = np.random.default_rng(0)
        a = ModelState({"iage": rng.uniform(1, 2, (nz, ny)), "iage_slow_rest": rng.uniform(1, 2, (nz, ny))})
        # log_vals / log in the reference's format: "<msg>[<module>,<region>]=<value>"
        with caplog.at_level(logging.INFO):
            a.log_vals("beta", a.norm())
            a.log("iterate")
        text = caplog.text
        assert "beta[iage,0]=" in text and "beta[iage,2]=" in text
        assert "iterate,mean[iage,1]=" in text and "iterate,norm[iage,2]=" in text
        # reversed operators
        b = a.__radd__(a)  # res = other + self
        np.testing.assert_allclose(b.get_tracer_vals("iage"), 2.0 * a.get_tracer_vals("iage"), rtol=1e-15)
        r = 2.0 / a
        np.testing.assert_allclose(r.get_tracer_vals("iage"), 2.0 / a.get_tracer_vals("iage"), rtol=1e-15)
        r = np.array([[1.0, 2.0, 3.0]]) / a  # [n_modules, region_cnt]: one scalar per column region
        np.testing.assert_allclose(r.get_tracer_vals("iage_slow_rest"),
                                   np.array([1.0, 2.0, 3.0])[None, :] / a.get_tracer_vals("iage_slow_rest"), rtol=1e-15)
        # the stats-variable methods of the surface, called the way newton_solver.py:52-58,330 calls them
        hist = str(tmp_path / "hist_00.nc")
        a.comp_fcn(None, None, hist)
        state = solver_state.SolverState("Newton", str(tmp_path))
        mods = [(t.name, None) for t in a.tracer_modules]
        stats = solver_state.StatsFile("Newton", str(tmp_path), 3, mods, solver_state.NEWTON_VARS)
        a.def_stats_vars(stats, hist, solver_state=state)
        a.put_stats_vars_iteration_invariant(stats, hist, solver_state=state)
        a.put_stats_vars(stats, hist, solver_state=state)
        assert state.step_logged("ModelStateBase.def_stats_vars", per_iteration=False)
        with netcdf_file(str(tmp_path / "Newton_stats.nc"), "r", mmap=False) as f:
            assert f.variables["iage"].dimensions == ("iteration", "depth", "ypos")
            assert f.variables["iage_mean_ypos"].dimensions == ("iteration", "depth")
            with netcdf_file(hist, "r", mmap=False) as h:
                vals = np.array(h.variables["iage"].data)
            w = np.full(61, 1.0 / 60.0)
            w[0] = w[-1] = 0.5 / 60.0
            np.testing.assert_allclose(np.array(f.variables["iage"].data)[0], np.einsum("i,i...", w, vals), rtol=1e-13)
            np.testing.assert_array_equal(np.array(f.variables["depth"].data), ModelState.depth.mid)
    finally:
        ModelState.reset()
