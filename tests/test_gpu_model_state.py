"""GPU parity tests through the host mirror of the reference's operator surface
(ModelState.comp_fcn / apply_precond_jacobian / comp_jacobian_fcn_state_prod / dot_prod ...)
against the reference's committed baselines (values in tests/golden/baselines.npz) with the
tolerances the reference's own CI scripts use."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def base(golden_dir):
    return np.load(os.path.join(golden_dir, "baselines.npz"))


def _modelinfo(tmp, nz, ny, vvel="0.1", kh="1000.0"):
    return {
        "model_name": "py_driver_2d", "tracer_module_names": "iage", "grid_vars_fname": os.path.join(tmp, "grid_vars.nc"),
        "depth_axisname": "depth", "depth_units": "m", "depth_edge_start": "0.0", "depth_edge_end": "4000.0",
        "depth_nlevs": str(nz), "depth_delta_ratio_max": "19.0",
        "ypos_axisname": "ypos", "ypos_units": "m", "ypos_edge_start": "0.0", "ypos_edge_end": "50.0e5",
        "ypos_nlevs": str(ny), "ypos_delta_ratio_max": "1.0", "max_abs_vvel": vvel, "horiz_mix_coeff": kh,
        "reinvoke": "False",
    }


def _state(ModelState, base, prefix):
    return ModelState({"iage": base[prefix + "/iage"], "iage_slow_rest": base[prefix + "/iage_slow_rest"]})


def _vals(ms):
    return np.stack([ms.get_tracer_vals("iage"), ms.get_tracer_vals("iage_slow_rest")])


def _want(base, prefix):
    return np.stack([base[prefix + "/iage"], base[prefix + "/iage_slow_rest"]])


def test_ci_py_driver_2d_iage(base, tmp_path):
    """scripts/ci_py_driver_2d_iage.sh: grid_vars.nc (default tol), fcn_0000/hist_0000/init_iterate
    (atol 1e-6, rtol 1e-3)"""
    from scipy.io import netcdf_file

    from nk_ooc_b200.py_driver_2d.model_state import ModelState
    from nk_ooc_b200.py_driver_2d.setup_solver import gen_grid_vars_file

    info = _modelinfo(str(tmp_path), 30, 30)
    gen_grid_vars_file(info)
    pre = "ci_py_driver_2d_iage/"
    with netcdf_file(info["grid_vars_fname"], "r", mmap=False) as f:
        for name in ("depth", "depth_edges", "depth_delta", "ypos", "ypos_edges", "ypos_delta", "grid_weight"):
            np.testing.assert_allclose(np.array(f.variables[name].data), base[pre + "grid_vars/" + name], rtol=1e-7, atol=2e-9)
        np.testing.assert_array_equal(np.array(f.variables["region_mask"].data), base[pre + "grid_vars/region_mask"])
    ModelState.configure(info)
    # gen_init_iterate == init_iterate_0000 of the baseline
    init = ModelState("gen_init_iterate")
    np.testing.assert_allclose(_vals(init), _want(base, pre + "init_iterate_0000"), rtol=1e-7, atol=2e-9)
    x = _state(ModelState, base, pre + "init_iterate_0000")
    fcn = x.comp_fcn(str(tmp_path / "fcn_0000.nc"), None, str(tmp_path / "hist_0000.nc"))
    np.testing.assert_allclose(_vals(fcn), _want(base, pre + "fcn_0000"), rtol=1e-3, atol=1e-6)
    # state file round trip
    back = ModelState(str(tmp_path / "fcn_0000.nc"))
    np.testing.assert_array_equal(_vals(back), _vals(fcn))
    # init_iterate = init_iterate_0000 + fcn (setup_solver.py:113-124)
    x += fcn
    np.testing.assert_allclose(_vals(x), _want(base, pre + "init_iterate"), rtol=1e-3, atol=1e-6)
    # hist file: 61 snapshots + process fields
    with netcdf_file(str(tmp_path / "hist_0000.nc"), "r", mmap=False) as f:
        for name in ("time", "stream", "vvel", "wvel", "horiz_mixing_coeff", "bldepth", "vert_mixing_coeff", "iage",
                     "iage_slow_rest", "iage_time_mean", "iage_time_std", "iage_time_delta", "iage_depth_int",
                     "iage_ypos_mean", "iage_depth_ypos_int"):
            np.testing.assert_allclose(np.array(f.variables[name].data), base[pre + "hist_0000/" + name], rtol=1e-3,
                                       atol=1e-6, err_msg=name)
    ModelState.reset()


def test_ci_py_driver_2d_iage_column_regions(base, tmp_path):
    """scripts/ci_py_driver_2d_iage_column_regions.sh: 20x3, no lateral processes -> 3 column
    regions; fcn_0000 (atol 1e-6/rtol 1e-3), precond_fcn_00 (rtol 2e-3), basis_00 (atol 5e-5),
    perturb_fcn_w_raw_00 (atol 5e-6)"""
    from nk_ooc_b200.py_driver_2d.model_state import ModelState
    from nk_ooc_b200.py_driver_2d.setup_solver import gen_grid_vars_file

    info = _modelinfo(str(tmp_path), 20, 3, "0.0", "0.0")
    gen_grid_vars_file(info)
    ModelState.configure(info)
    pre = "ci_py_driver_2d_iage_column_regions/"
    cfg = ModelState.model_config_obj
    assert cfg.region_cnt == 3
    np.testing.assert_array_equal(cfg.region_mask, base[pre + "grid_vars/region_mask"])
    x0 = _state(ModelState, base, pre + "init_iterate_0000")
    fcn0 = x0.comp_fcn(None, None)
    np.testing.assert_allclose(_vals(fcn0), _want(base, pre + "fcn_0000"), rtol=1e-3, atol=1e-6)
    # Newton iteration 0: iterate_00 = init_iterate (spin-up result)
    iterate = _state(ModelState, base, pre + "init_iterate")
    fcn = iterate.comp_fcn(None, None, str(tmp_path / "hist_00.nc"))
    assert fcn.norm().shape == (1, 3)
    iterate.gen_precond_jacobian(str(tmp_path / "hist_00.nc"), str(tmp_path / "precond_00.nc"))
    precond_fcn = fcn.apply_precond_jacobian(str(tmp_path / "precond_00.nc"), None, None)
    np.testing.assert_allclose(_vals(precond_fcn), _want(base, pre + "precond_fcn_00"), rtol=2e-3, atol=1e-9)
    beta = precond_fcn.norm()
    basis = -precond_fcn / beta
    np.testing.assert_allclose(_vals(basis), _want(base, pre + "basis_00"), rtol=1e-7, atol=5e-5)
    np.testing.assert_allclose(basis.norm(), 1.0, rtol=1e-12)
    # FD Jacobian-vector product with the baseline's basis vector
    direction = _state(ModelState, base, pre + "basis_00")
    w_raw = iterate.comp_jacobian_fcn_state_prod(fcn, direction, None, None)
    sigma = 1.0e-4 * iterate.norm()
    perturb_fcn = w_raw * sigma + fcn
    np.testing.assert_allclose(_vals(perturb_fcn), _want(base, pre + "perturb_fcn_w_raw_00"), rtol=1e-7, atol=5e-6)
    ModelState.reset()


def test_batched_members_equal_single_states(tmp_path):
    """B members evaluated at once == the same states evaluated one by one; operators broadcast"""
    from nk_ooc_b200.py_driver_2d.model_state import ModelState
    from nk_ooc_b200.py_driver_2d.setup_solver import gen_grid_vars_file

    info = _modelinfo(str(tmp_path), 12, 8)
    gen_grid_vars_file(info)
    ModelState.configure(info, steps_per_year=60)
    rng = np.random.default_rng(2)
    singles = []
    for _ in range(5):
        singles.append(ModelState({"iage": rng.normal(size=(12, 8)), "iage_slow_rest": rng.normal(size=(12, 8))}))
    batch = ModelState.from_members(singles)
    fb = batch.comp_fcn(None, None)
    assert fb.norm().shape == (1, 1, 5)
    for b, s in enumerate(singles):
        fs = s.comp_fcn(None, None)
        np.testing.assert_allclose(_vals(fb.member(b)), _vals(fs), rtol=0, atol=1e-13 * np.abs(_vals(fs)).max())
        np.testing.assert_allclose(fb.norm()[0, 0, b], fs.norm()[0, 0], rtol=1e-13)
    # modified Gram-Schmidt against an orthonormalised pair of in-memory basis vectors
    v0 = singles[0] / singles[0].norm()
    w = singles[1]._like()
    h = w.mod_gram_schmidt(1, lambda q, i: v0, "basis")
    assert h.shape == (1, 1, 1)
    np.testing.assert_allclose(w.dot_prod(v0), 0.0, atol=1e-13)
    ModelState.reset()


@pytest.mark.parametrize("B", [1, 4])
def test_py_driver_2d_phosphorus_preconditioner(golden_dir, tmp_path, B):
    """ModelState.apply_precond_jacobian for py_driver_2d phosphorus (phosphorus.py:197-274).
    (i) device path (banded LU pair in tracer-fastest ordering, K5/K6 kernels) against a dense
    host solve with the SAME shift and null vector: rounding level; (ii) against the reference's
    own result, at the level to which the reference reproduces itself (ARPACK's second eigenvalue at
    sigma = 0 scatters by ~1e-3 between calls, see tests/test_oracle.py): 3e-2 of the maximum."""
    from scipy.io import netcdf_file

    from nk_ooc_b200.py_driver_2d.model_state import ModelState
    from nk_ooc_b200.py_driver_2d.setup_solver import gen_grid_vars_file

    g = np.load(os.path.join(golden_dir, "py_driver_2d.npz"))
    tag = "g14x11"
    nz, ny = 14, 11
    info = _modelinfo(str(tmp_path), nz, ny)
    info["tracer_module_names"] = "phosphorus"
    gen_grid_vars_file(info)
    ModelState.configure(info)
    y = g[f"{tag}/phosphorus/precond_y"]
    precond_fname = str(tmp_path / "precond_00.nc")
    with netcdf_file(precond_fname, "w", version=2) as f:
        f.createDimension("time", None)
        f.createDimension("depth", nz)
        f.createDimension("ypos", ny)
        f.createVariable("time", "f8", ("time",))
        f.createVariable("po4", "f8", ("time", "depth", "ypos"))
        f.variables["time"][:] = g[f"{tag}/phosphorus/precond_times"]
        f.variables["po4"][:] = g[f"{tag}/phosphorus/precond_po4"]
    rng = np.random.default_rng(17)
    ys = [y] + [rng.normal(size=y.shape) * np.array([1.0, 0.05, 0.01])[:, None, None] for _ in range(B - 1)]
    ms = ModelState("zeros", members=B)
    tms = ms.tracer_modules[0]
    for b, yb in enumerate(ys):
        tms.vals[..., b] = torch.from_numpy(yb).cuda()
    res = ms.apply_precond_jacobian(precond_fname, None, None)
    got = res.tracer_modules[0].vals[..., :B].cpu().numpy()
    want = g[f"{tag}/phosphorus/precond"]
    np.testing.assert_allclose(got[..., 0], want, rtol=0, atol=3e-2 * np.abs(want).max())
    # same shift and null vector, dense float64 solve on the host
    _, _, _, e_vect, shift = ModelState._precond_cache[(tms.name, precond_fname)]
    from oracle import nk_oracle as o

    grid = o.Grid2D(g[f"{tag}/depth_edges"], g[f"{tag}/ypos_edges"], 0.1, 1000.0)
    tv = np.zeros((3, nz, ny))
    tv[0] = g[f"{tag}/phosphorus/precond_po4"][-1]
    T = 365.0 * 86400.0
    mat = T * o.Phosphorus2D(grid).comp_jacobian(0.5 * T, tv.reshape(-1)).toarray()
    weight = np.outer(grid.depth.delta, grid.ypos.delta)
    weight = weight / weight.sum()
    eye = np.eye(mat.shape[0])
    for b, yb in enumerate(ys):
        sol = 2.0 * np.linalg.solve(mat - 0.5 * shift * eye, yb.reshape(-1)) - np.linalg.solve(mat - shift * eye, yb.reshape(-1))
        sol = sol.reshape(3, nz, ny)
        sol = sol - (weight[None] * sol).sum() * e_vect
        ref = sol - yb
        np.testing.assert_allclose(got[..., b], ref, rtol=0, atol=1e-7 * np.abs(ref).max())
    ModelState.reset()
