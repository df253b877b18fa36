"""precond matrix definitions of a tracer_module_defs file (nk_ooc/model_config.py:197-246): the same checks as the
reference's tests/test_model_config.py:25-57, on the definitions its input files hold (restated here)"""
import copy
import os

import pytest

from nk_ooc_b200.model_state_base import (_expand_matrix_defs, check_precond_matrix_defs,
                                          propagate_base_matrix_defs_to_all)

# input/test_problem/tracer_module_defs.yaml:57-64 and input/py_driver_2d/tracer_module_defs.yaml:53-62
TEST_PROBLEM = {
    "base": {"hist_to_precond_varnames": ["mixing_coeff:mean", "mixing_coeff:log_mean"]},
    "phosphorus": {"hist_to_precond_varnames": ["po4_s_restore_tau_r:mean"]},
}
PY_DRIVER_2D = {
    "base": {"hist_to_precond_varnames": ["time"]},
    "phosphorus": {"hist_to_precond_varnames": ["po4"]},
    "forced_{suff}": {"hist_to_precond_varnames": ["{suff}"]},
}


def test_propagate_base_matrix_defs_to_all():
    defs = copy.deepcopy(TEST_PROBLEM)
    propagate_base_matrix_defs_to_all(defs)
    base, phosphorus = defs["base"], defs["phosphorus"]
    assert phosphorus["hist_to_precond_varnames"] == ["po4_s_restore_tau_r:mean", "mixing_coeff:mean",
                                                      "mixing_coeff:log_mean"]
    # a hist variable added to base reaches phosphorus on the next propagation
    base["hist_to_precond_varnames"].append("new_hist_var")
    propagate_base_matrix_defs_to_all(defs)
    assert "new_hist_var" in phosphorus["hist_to_precond_varnames"]
    # a key phosphorus lacks is taken over (as a copy)
    base["precond_matrices_opts"] = ["matrix_opt_A sub_opt"]
    propagate_base_matrix_defs_to_all(defs)
    assert phosphorus["precond_matrices_opts"] == ["matrix_opt_A sub_opt"]
    assert phosphorus["precond_matrices_opts"] is not base["precond_matrices_opts"]
    # an option both set keeps the matrix's own sub-option; nothing is added twice
    base["precond_matrices_opts"].append("matrix_opt_B sub_opt_base")
    phosphorus["precond_matrices_opts"].append("matrix_opt_B sub_opt_phosphorus")
    propagate_base_matrix_defs_to_all(defs)
    assert "matrix_opt_B sub_opt_phosphorus" in phosphorus["precond_matrices_opts"]
    assert "matrix_opt_B sub_opt_base" not in phosphorus["precond_matrices_opts"]
    assert phosphorus["precond_matrices_opts"].count("matrix_opt_A sub_opt") == 1
    # dict settings gain the missing keys only; other types are refused
    base["d"] = {"a": 1, "b": 2}
    phosphorus["d"] = {"a": 10}
    propagate_base_matrix_defs_to_all(defs)
    assert phosphorus["d"] == {"a": 10, "b": 2}
    base["n"], phosphorus["n"] = 1, 2
    with pytest.raises(TypeError):
        propagate_base_matrix_defs_to_all(defs)
    # no base: nothing happens
    lone = {"m": {"hist_to_precond_varnames": ["x"]}}
    propagate_base_matrix_defs_to_all(lone)
    assert lone == {"m": {"hist_to_precond_varnames": ["x"]}}


def test_suffix_expansion_and_vetting():
    defs = _expand_matrix_defs(PY_DRIVER_2D, ["iage", "forced_{suff}:o2_like:dye"])
    propagate_base_matrix_defs_to_all(defs)
    check_precond_matrix_defs(defs)
    assert list(defs) == ["base", "phosphorus", "forced_o2_like", "forced_dye"]
    assert defs["forced_o2_like"]["hist_to_precond_varnames"] == ["o2_like", "time"]
    assert defs["forced_dye"]["hist_to_precond_varnames"] == ["dye", "time"]
    assert "forced_{suff}" in PY_DRIVER_2D and PY_DRIVER_2D["forced_{suff}"]["hist_to_precond_varnames"] == ["{suff}"]
    defs["phosphorus"]["hist_to_precond_varnames"].append("po4:median")
    with pytest.raises(ValueError, match="unknown time_op=median"):
        check_precond_matrix_defs(defs)


def _gen_precond(hist_vars, hist_fname, precond_fname):
    """ModelStateBase.gen_precond_jacobian needs nothing of a state but its list of hist variables"""
    from types import SimpleNamespace

    from nk_ooc_b200.model_state_base import ModelStateBase

    stub = SimpleNamespace(hist_vars_for_precond_list=lambda: list(hist_vars))
    ModelStateBase.gen_precond_jacobian(stub, hist_fname, precond_fname)


def test_precond_file_from_the_baselines_hist_files(tmp_path):
    """hist file -> precond file (model_state_base.py:404-481): from the reference's OWN hist files the result has the
    metadata (dimension and variable order, names, attributes) and the values of the reference's precond files"""
    from baseline_files import materialise
    from nk_ooc_b200 import baseline_cmp

    base = materialise(str(tmp_path / "baselines"))
    # py_driver_2d iage: only `time` (input/py_driver_2d/tracer_module_defs.yaml:53-56)
    cfg = "ci_py_driver_2d_iage_column_regions"
    out = tmp_path / "a"
    _gen_precond(["time"], os.path.join(base, cfg, "hist_0000.nc"), str(out / "precond_00.nc"))
    assert baseline_cmp.compare("precond_00.nc", str(out), os.path.join(base, cfg))
    # test_problem iage: the two reductions of the mixing coefficient, which does not depend on the state — the hist
    # file of ci_short's first iterate gives the precond file of ci_long_iage
    out = tmp_path / "b"
    _gen_precond(["mixing_coeff:mean", "mixing_coeff:log_mean"], os.path.join(base, "ci_short", "hist_00.nc"),
                 str(out / "precond_00.nc"))
    assert baseline_cmp.compare("precond_00.nc", str(out), os.path.join(base, "ci_long_iage"))
    # with the phosphorus matrix (input/test_problem/tracer_module_defs.yaml:62-64): dimensions and coordinates of all
    # results first, then the results in the order of the list
    from scipy.io import netcdf_file

    out = tmp_path / "c"
    _gen_precond(["po4_s_restore_tau_r:mean", "mixing_coeff:mean", "mixing_coeff:log_mean"],
                 os.path.join(base, "ci_short", "hist_00.nc"), str(out / "precond_00.nc"))
    with netcdf_file(str(out / "precond_00.nc"), "r", mmap=False) as got, \
            netcdf_file(os.path.join(base, "ci_short", "hist_00.nc"), "r", mmap=False) as hist:
        assert list(got.dimensions) == ["depth", "depth_edges"]
        # (scipy's writer orders the variables of a file by shape, largest first, creation order within a shape)
        assert list(got.variables) == ["depth_edges", "mixing_coeff_mean", "mixing_coeff_log_mean", "depth",
                                       "po4_s_restore_tau_r_mean"]
        tau = hist.variables["po4_s_restore_tau_r"]
        res = got.variables["po4_s_restore_tau_r_mean"]
        assert res.dimensions == ("depth",)
        assert res.long_name.decode() == tau.long_name.decode() + ", mean over time dim"
        assert not hasattr(res, "cell_methods") or b"time:" not in res.cell_methods
        import numpy as np

        np.testing.assert_allclose(res.data, np.array(tau.data).mean(axis=0), rtol=1e-15)
