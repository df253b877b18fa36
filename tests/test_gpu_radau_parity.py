"""GPU parity of F(x) = x(T) - x(0) against the REFERENCE's own integrator.

Truth: tests/golden/radau_<grid>_<module>.npz, produced by oracle/gen_golden_radau.py from the
reference's own tracer-module classes and its own solve_ivp(Radau) call
(nk_ooc/py_driver_2d/model_state.py:102-114, forced.py:114-154, phosphorus.py:58-95,
test_problem/model_state.py:83-92, dye_decay.py:26-47) at rtol = atol = 1e-9 (1e-12 for test_problem),
plus the same run at the reference's own tolerance 1e-6 so that the reference's own error is on record.

Stated tolerance: the reference's CI tolerance for py_driver_2d function
evaluations, scripts/ci_py_driver_2d_iage.sh:25-41 — rtol 1e-3, atol 1e-6 — with atol multiplied by the
tracer's own scale max|x0| (the CI's iage fields are O(1..100) years; po4 is O(1), dop / pop O(1e-2)):

    |F_gpu - F_radau(1e-9)| <= 1e-3 |F_radau| + 1e-6 max(1, max|x0_tracer|)      (FCN_RTOL, FCN_ATOL)

The product's default (graded, 2640 steps / yr) schedule is what is tested; the error of the reference
at ITS tolerance against the same truth is asserted to be of the same order, so the GPU path is as close
to the truth as the reference itself is."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

FCN_RTOL = 1.0e-3
FCN_ATOL = 1.0e-6
YEAR = 365.0 * 86400.0


def _load(golden_dir, grid, module):
    import radau_cases

    g = radau_cases.load(grid, module)
    if g is None:
        pytest.skip(f"radau_{grid}_{module}.npz not generated")
    return g


def _model(g, module):
    import radau_cases

    return radau_cases.model(g, module)


def _eval(model, x0, hist_idx=None):
    from nk_ooc_b200.engine import padded_members

    B = 3  # the state, and two copies (a batch exercises the member-fastest layout)
    xd = torch.zeros(x0.shape + (padded_members(B),), dtype=torch.float64, device="cuda")
    xd[..., :B] = torch.from_numpy(x0).cuda()[..., None]
    if hist_idx is None:
        f = model.eval(xd, B)
        snaps = None
    else:
        times = np.linspace(0.0, YEAR, 61)[hist_idx]
        f, snaps = model.eval(xd, B, hist_steps=model.step_index_of_times(times))
        snaps = snaps.cpu().numpy()
    torch.cuda.synchronize()
    model.check_health()
    f = f.cpu().numpy()
    assert np.array_equal(f[..., 0], f[..., 1]) and np.array_equal(f[..., 0], f[..., 2])
    return f[..., 0], snaps


def _tol_ratio(got, want, x0):
    """worst |got - want| / (rtol |want| + atol scale_tracer) over the field"""
    scale = np.maximum(1.0, np.abs(x0).reshape(x0.shape[0], -1).max(axis=1))[:, None, None]
    return float((np.abs(got - want) / (FCN_RTOL * np.abs(want) + FCN_ATOL * scale)).max())


@pytest.mark.parametrize("grid", ["g14x11", "g30x30", "g40x50", "g80x100", "g125x150"])
@pytest.mark.parametrize("module", ["iage", "forced", "phosphorus"])
def test_fcn_vs_reference_radau(golden_dir, grid, module):
    from nk_ooc_b200.py_driver_2d.model_state import default_schedule, hist_schedule

    g = _load(golden_dir, grid, module)
    model = _model(g, module)
    nz = int(g["params"][0])
    model.set_graded_schedule(**default_schedule(module, nz))
    x0 = g["x0"]
    truth = g["tol1e-09/fcn"]
    idx = [int(i) for i in g["tol1e-09/snap_idx"]]
    got, snaps = _eval(model, x0, idx)
    if hist_schedule(module, nz) != default_schedule(module, nz):
        # (phosphorus on fine grids: the evaluations that write a hist file take the snapshots from a second integration
        # with the mixed-layer ramps resolved more finely, F from the schedule of every other evaluation)
        model.set_graded_schedule(**hist_schedule(module, nz))
        _, snaps = _eval(model, x0, idx)
    ratio = _tol_ratio(got, truth, x0)
    ref_ratio = _tol_ratio(g["tol1e-06/fcn"], truth, x0) if "tol1e-06/fcn" in g.files else float("nan")
    print(f"{grid}/{module}: max|F| {np.abs(truth).max():.3e}  max|dF| gpu {np.abs(got - truth).max():.3e} "
          f"(ratio to tolerance {ratio:.3f}); the reference at 1e-6: {ref_ratio:.3f}")
    assert ratio <= 1.0
    # the hist snapshots the reference would write (x(t_k), py_driver_2d/model_state.py:80-83)
    want_snaps = g["tol1e-09/snaps"]
    for i in range(len(idx)):
        assert _tol_ratio(snaps[i], want_snaps[i], x0) <= 1.0, f"snapshot {idx[i]}"
    if module == "forced":
        # the sink limiter of forced.py:140-152 must have been active AND inactive somewhere
        q = x0[0] / 0.05
        assert ((q > 0) & (q < 1)).any() and (q > 1).any()


def test_forced_model_state_comp_fcn_vs_reference_radau(golden_dir, tmp_path):
    """the same comparison through the host mirror of the operator surface: ModelState.comp_fcn with the
    o2_like options of scripts/run_py_driver_2d_forced_o2_like.sh read from modelinfo and the sink record
    read from a netCDF file (the mirror of utils.gen_forcing_fcn, here on the model grid already)"""
    from scipy.io import netcdf_file

    from nk_ooc_b200.py_driver_2d.model_state import ModelState
    from nk_ooc_b200.py_driver_2d.setup_solver import gen_grid_vars_file

    g = _load(golden_dir, "g14x11", "forced")
    nz, ny = int(g["params"][0]), int(g["params"][1])
    fname = str(tmp_path / "sms.nc")
    depth_mid = 0.5 * (g["depth_edges"][1:] + g["depth_edges"][:-1])
    ypos_mid = 0.5 * (g["ypos_edges"][1:] + g["ypos_edges"][:-1])
    with netcdf_file(fname, "w", version=2) as f:
        f.createDimension("time", len(g["frc_time"]))
        f.createDimension("depth", nz)
        f.createDimension("ypos", ny)
        for name, vals in (("time", g["frc_time"]), ("depth", depth_mid), ("ypos", ypos_mid)):
            f.createVariable(name, "f8", (name,))[:] = vals
        # the file holds the record BEFORE scalef; modelinfo carries forced_sms_scalef
        f.createVariable("po4_sms", "f8", ("time", "depth", "ypos"))[:] = -3.0 * g["frc_data"]
    info = {
        "model_name": "py_driver_2d", "tracer_module_names": "forced_{suff}:o2_like",
        "grid_vars_fname": str(tmp_path / "grid_vars.nc"),
        "depth_axisname": "depth", "depth_units": "m", "depth_edge_start": "0.0", "depth_edge_end": "4000.0",
        "depth_nlevs": str(nz), "depth_delta_ratio_max": "19.0",
        "ypos_axisname": "ypos", "ypos_units": "m", "ypos_edge_start": "0.0", "ypos_edge_end": "50.0e5",
        "ypos_nlevs": str(ny), "ypos_delta_ratio_max": "1.0", "max_abs_vvel": "0.1", "horiz_mix_coeff": "1000.0",
        "reinvoke": "False",
        "forced_surf_restore_opt": "const", "forced_surf_restore_const": "1.0",
        "forced_surf_restore_rate_10m": "1.0 / 3600.0", "forced_sms_opt": "file", "forced_sms_fname": fname,
        "forced_sms_varname": "po4_sms", "forced_sms_scalef": "-1.0 / 3.0", "forced_sink_thres": "0.05",
    }
    gen_grid_vars_file(info)
    ModelState.configure(info)
    try:
        x = ModelState({"o2_like": g["x0"][0]})
        fcn = x.comp_fcn(None, None)
        got = fcn.get_tracer_vals("o2_like")[None]
        assert _tol_ratio(got, g["tol1e-09/fcn"], g["x0"]) <= 1.0
    finally:
        ModelState.reset()


def test_phosphorus_fine_grid_hist_file_vs_reference_radau(golden_dir, tmp_path):
    """ModelState.comp_fcn WITH a hist file on 80 x 100: F is the F of every other evaluation (bit-identical with the
    evaluation without a hist file: the finite-difference products difference the two), the snapshots in the hist file
    come from the second, finer integration (model_state.py:hist_schedule) and meet the tolerance against the
    reference's Radau solution, which the 2640-step snapshots do not in the middle of the first mixed-layer ramp"""
    from scipy.io import netcdf_file

    from nk_ooc_b200.py_driver_2d.model_state import ModelState, default_schedule, hist_schedule
    from nk_ooc_b200.py_driver_2d.setup_solver import gen_grid_vars_file

    g = _load(golden_dir, "g80x100", "phosphorus")
    nz, ny, ratio = int(g["params"][0]), int(g["params"][1]), float(g["params"][2])
    assert hist_schedule("phosphorus", nz) != default_schedule("phosphorus", nz)
    info = {
        "model_name": "py_driver_2d", "tracer_module_names": "phosphorus",
        "grid_vars_fname": str(tmp_path / "grid_vars.nc"),
        "depth_axisname": "depth", "depth_units": "m", "depth_edge_start": "0.0", "depth_edge_end": "4000.0",
        "depth_nlevs": str(nz), "depth_delta_ratio_max": repr(ratio),
        "ypos_axisname": "ypos", "ypos_units": "m", "ypos_edge_start": "0.0", "ypos_edge_end": "50.0e5",
        "ypos_nlevs": str(ny), "ypos_delta_ratio_max": "1.0", "max_abs_vvel": repr(float(g["params"][3])),
        "horiz_mix_coeff": repr(float(g["params"][4])), "reinvoke": "False",
    }
    gen_grid_vars_file(info)
    ModelState.configure(info)
    try:
        np.testing.assert_allclose(ModelState.depth.edges, g["depth_edges"], rtol=1e-14)
        names = ("po4", "dop", "pop")
        x = ModelState({name: g["x0"][i] for i, name in enumerate(names)})
        plain = x.comp_fcn(None, None)
        hist_fname = str(tmp_path / "hist.nc")
        with_hist = x.comp_fcn(None, None, hist_fname)
        for name in names:
            assert np.array_equal(plain.get_tracer_vals(name), with_hist.get_tracer_vals(name)), name
        got = np.stack([plain.get_tracer_vals(name) for name in names])
        assert _tol_ratio(got, g["tol1e-09/fcn"], g["x0"]) <= 1.0
        idx = [int(i) for i in g["tol1e-09/snap_idx"]]
        with netcdf_file(hist_fname, "r", mmap=False) as nc:
            snaps = np.stack([np.array(nc.variables[name].data)[idx] for name in names], axis=1)
        worst = max(_tol_ratio(snaps[i], g["tol1e-09/snaps"][i], g["x0"]) for i in range(len(idx)))
        print(f"hist snapshots {idx}: worst ratio to the tolerance {worst:.3f}")
        assert worst <= 1.0
    finally:
        ModelState.reset()


def test_test_problem_dye_decay_fcn_vs_reference_radau(golden_dir, tmp_path):
    """test_problem dye_decay F(x) (both parameterised modules of scripts/ci_long_dye_decay.sh) through
    ModelState.comp_fcn against the reference's own call (test_problem/model_state.py:83-92, rtol = atol =
    1e-12) at the CI's default comparison tolerance rtol 1e-7 / atol 2e-9 (nk_ooc/baseline_cmp.py:20-25) —
    ci_long_dye_decay.sh itself pins only Newton_state.json"""
    from nk_ooc_b200.spatial_axis import spatial_axis_from_defn
    from nk_ooc_b200.test_problem.model_state import ModelState, gen_depth_axis_file

    g = {name: _load(golden_dir, "tp20", name) for name in ("dye_decay_001", "dye_decay_010")}
    info = {"model_name": "test_problem", "tracer_module_names": "dye_decay_{suff}:001:010",
            "grid_vars_fname": str(tmp_path / "depth_axis.nc"), "depth_axisname": "depth", "reinvoke": "False"}
    depth = spatial_axis_from_defn("depth", nlevs=20)
    np.testing.assert_array_equal(depth.edges, g["dye_decay_010"]["depth_edges"])
    gen_depth_axis_file(info, depth)
    ModelState.configure(info)
    try:
        x = ModelState({name: g[name]["x0"][0] for name in g})
        fcn = x.comp_fcn(None, None)
        for name in g:
            np.testing.assert_allclose(fcn.get_tracer_vals(name), g[name]["fcn"][0], rtol=1e-7, atol=2e-9, err_msg=name)
    finally:
        ModelState.reset()
