"""GPU tests of the device-resident Newton-Krylov driver (nk_ooc_b200/solver.py) against the
reference's committed baselines of a complete Newton step (krylov_res_00, increment_00,
iterate_01) with the tolerances of the reference's own CI scripts, and to convergence."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from test_gpu_model_state import _modelinfo, _state, _vals, _want  # noqa: E402
from test_gpu_test_problem import _configure  # noqa: E402


@pytest.fixture(scope="module")
def base(golden_dir):
    return np.load(os.path.join(golden_dir, "baselines.npz"))


TP_SOLVERINFO = {"newton_rel_tol": "1.0e-8", "newton_max_iter": "5", "post_newton_fp_iter": "1",
                 "krylov_rel_tol": "0.01"}  # input/test_problem/newton_krylov.cfg:32-43
PD_SOLVERINFO = {"newton_rel_tol": "1.0e-5", "newton_max_iter": "5", "post_newton_fp_iter": "1",
                 "krylov_rel_tol": "0.01"}  # input/py_driver_2d/newton_krylov.cfg:32-43


def test_ci_long_iage_newton_step_and_convergence(base, tmp_path):
    """scripts/ci_long_iage.sh: krylov_res_00, increment_00, iterate_01 at rtol 2e-4; the reference run
    stops at Newton iteration 3 (baselines/ci_long_iage/Newton_state.json)"""
    from scipy.io import netcdf_file

    from nk_ooc_b200.solver import NewtonSolver

    ModelState = _configure(str(tmp_path), "iage")
    pre = "ci_long_iage/"
    iterate = ModelState({"iage": base["ci_short/init_iterate/iage"]})
    solver = NewtonSolver(iterate, TP_SOLVERINFO, workdir=str(tmp_path / "work"))
    assert not solver.converged_flat()
    increment = solver.step()
    np.testing.assert_allclose(increment.get_tracer_vals("iage"), base[pre + "increment_00/iage"], rtol=2e-4, atol=2e-9)
    np.testing.assert_allclose(solver.iterate.get_tracer_vals("iage"), base[pre + "iterate_01/iage"], rtol=2e-4,
                               atol=2e-9)
    # the files the reference's CI compares are written with the reference's names
    with netcdf_file(str(tmp_path / "work" / "krylov_00" / "krylov_res_00.nc"), "r", mmap=False) as f:
        np.testing.assert_allclose(np.array(f.variables["iage"].data), base[pre + "krylov_res_00/iage"], rtol=2e-4,
                                   atol=2e-9)
    for name in ("increment_00.nc", "iterate_01.nc", "fcn_01.nc", "hist_01.nc"):
        assert os.path.exists(str(tmp_path / "work" / name)), name
    rec = solver.history[-1]
    assert rec["krylov_iterations"] >= 1 and (rec["armijo_factor"] == 1.0).all() and rec["armijo_ind"] == 0
    assert (rec["krylov_precond_resid_norm"][-1] < 0.01 * rec["krylov_beta"]).all()
    solver.solve()
    assert solver.converged_flat() and solver.iteration == 3
    norms = [float(r["fcn_norm"][0, 0]) for r in solver.history]
    assert all(b < a for a, b in zip(norms[:-1], norms[1:])), norms
    ModelState.reset()


def test_speculative_armijo_equals_sequential(base, tmp_path):
    """armijo_batch = 3 (three candidates as members of one batched evaluation) gives the
    sequential result; a deliberately overshooting increment exercises the halving"""
    from nk_ooc_b200.solver import NewtonSolver

    ModelState = _configure(str(tmp_path), "iage")
    x0 = ModelState({"iage": base["ci_short/init_iterate_00/iage"]})
    out = []
    for k in (1, 3):
        solver = NewtonSolver(x0._like(), TP_SOLVERINFO, workdir=str(tmp_path / f"w{k}"), armijo_batch=k, dump=False)
        inc = solver.fcn * 40.0  # far too long a step along F: needs damping
        prov, prov_fcn, factor, ind = solver._comp_next_iterate(inc)
        out.append((prov.get_tracer_vals("iage"), prov_fcn.get_tracer_vals("iage"), factor, ind))
    assert out[0][3] == out[1][3] and out[0][3] >= 1
    np.testing.assert_array_equal(out[0][2], out[1][2])
    np.testing.assert_allclose(out[0][0], out[1][0], rtol=0, atol=1e-13 * np.abs(out[0][0]).max())
    np.testing.assert_allclose(out[0][1], out[1][1], rtol=0, atol=1e-12 * np.abs(out[0][1]).max())
    # the same with the work directory kept: the halvings are saved step by step (armijo_ind, armijo_factor), the
    # accepted candidate and its function value are files, and a second call — "_comp_next_iterate complete" is
    # logged — reads them back instead of evaluating again (newton_solver.py:199-206)
    import json

    from nk_ooc_b200 import _lib

    work = str(tmp_path / "wd")
    solver = NewtonSolver(x0._like(), TP_SOLVERINFO, workdir=work)
    inc = solver.fcn * 40.0
    prov, prov_fcn, factor, ind = solver._comp_next_iterate(inc)
    assert ind == out[0][3]
    np.testing.assert_array_equal(factor, out[0][2])
    state = json.load(open(os.path.join(work, "Newton_state.json")))
    assert state["armijo_ind"] == ind
    assert f"00:comp_fcn complete for {work}/prov_fcn_Armijo_{ind:02}_00.nc" in state["step_log"]
    assert "00:_comp_next_iterate complete" in state["step_log"]
    assert os.path.exists(os.path.join(work, f"prov_Armijo_{ind:02}_00.nc"))
    assert not os.path.exists(os.path.join(work, f"prov_hist_Armijo_{ind - 1:02}_00.nc"))  # only the latest hist file stays
    lib = _lib.load()
    n0 = lib.nkb_launch_count()
    again, again_fcn, factor2, ind2 = solver._comp_next_iterate(inc)
    assert lib.nkb_launch_count() == n0 and ind2 == ind
    np.testing.assert_array_equal(again.get_tracer_vals("iage"), prov.get_tracer_vals("iage"))
    np.testing.assert_array_equal(again_fcn.get_tracer_vals("iage"), prov_fcn.get_tracer_vals("iage"))
    ModelState.reset()


def test_ci_py_driver_2d_iage_column_regions_newton_step(base, tmp_path):
    """scripts/ci_py_driver_2d_iage_column_regions.sh: krylov_res_00, increment_00, iterate_01 at
    rtol 1.9e-2 (3 column regions: per-region Krylov coefficients and Armijo factors)"""
    from scipy.io import netcdf_file

    from nk_ooc_b200.py_driver_2d.model_state import ModelState
    from nk_ooc_b200.py_driver_2d.setup_solver import gen_grid_vars_file
    from nk_ooc_b200.solver import NewtonSolver

    info = _modelinfo(str(tmp_path), 20, 3, "0.0", "0.0")
    gen_grid_vars_file(info)
    ModelState.configure(info)
    pre = "ci_py_driver_2d_iage_column_regions/"
    iterate = _state(ModelState, base, pre + "init_iterate")
    solver = NewtonSolver(iterate, PD_SOLVERINFO, workdir=str(tmp_path / "work"))
    increment = solver.step()
    np.testing.assert_allclose(_vals(increment), _want(base, pre + "increment_00"), rtol=1.9e-2, atol=1e-9)
    np.testing.assert_allclose(_vals(solver.iterate), _want(base, pre + "iterate_01"), rtol=1.9e-2, atol=1e-9)
    with netcdf_file(str(tmp_path / "work" / "krylov_00" / "krylov_res_00.nc"), "r", mmap=False) as f:
        got = np.stack([np.array(f.variables[n].data) for n in ("iage", "iage_slow_rest")])
    np.testing.assert_allclose(got, _want(base, pre + "krylov_res_00"), rtol=1.9e-2, atol=1e-9)
    rec = solver.history[-1]
    assert rec["krylov_beta"].shape == (1, 3) and (rec["armijo_factor"] == 1.0).all()
    ModelState.reset()


def test_coloured_column_probes_give_the_jacobian(base, tmp_path):
    """all (tracer, level) unit perturbations of all columns in ONE batched evaluation
    (3 colours x 2 tracers x 20 levels = 120 members, nk_ooc_b200/colouring.py): the decoded
    column blocks reproduce the finite-difference Jacobian-vector product of
    comp_jacobian_fcn_state_prod; without lateral processes the off-column blocks vanish"""
    from nk_ooc_b200 import colouring as col
    from nk_ooc_b200.py_driver_2d.model_state import ModelState
    from nk_ooc_b200.py_driver_2d.setup_solver import gen_grid_vars_file

    info = _modelinfo(str(tmp_path), 20, 3, "0.0", "0.0")
    gen_grid_vars_file(info)
    ModelState.configure(info)
    pre = "ci_py_driver_2d_iage_column_regions/"
    iterate = _state(ModelState, base, pre + "init_iterate")
    x0 = _vals(iterate)  # [T, nz, ny]
    T, nz, ny = x0.shape
    colour = col.column_colouring(ny)
    eps = 1.0e-2
    probes = col.probe_batch(x0, colour, eps)
    B = probes.shape[0]
    assert B == 3 * T * nz
    batched = ModelState("zeros", members=B)
    tms = batched.tracer_modules[0]
    tms.vals[..., :B] = torch.from_numpy(np.ascontiguousarray(np.moveaxis(probes, 0, -1))).cuda()
    fb = batched.comp_fcn(None, None)
    fprobe = np.moveaxis(fb.tracer_modules[0].vals[..., :B].cpu().numpy(), -1, 0)
    f0 = _vals(iterate.comp_fcn(None, None))
    jac = col.decode_probes(f0, fprobe, colour, eps, reach=1)
    scale = np.abs(jac[:, 1]).max()
    assert np.abs(jac[:, 0]).max() <= 1e-9 * scale and np.abs(jac[:, 2]).max() <= 1e-9 * scale
    rng = np.random.default_rng(2)
    v = rng.normal(size=x0.shape)
    direction = ModelState({"iage": v[0], "iage_slow_rest": v[1]})
    fcn = iterate.comp_fcn(None, None)
    jv = _vals(iterate.comp_jacobian_fcn_state_prod(fcn, direction, None, None))
    want = np.stack([(jac[j, 1] @ v[:, :, j].reshape(-1)).reshape(T, nz) for j in range(ny)], axis=-1)
    np.testing.assert_allclose(jv, want, rtol=0, atol=1e-6 * np.abs(want).max())
    ModelState.reset()


def test_ci_long_dye_decay_newton_convergence(tmp_path):
    """scripts/ci_long_dye_decay.sh: two parameterised modules (dye_decay_{suff}:001:010), init iterate
    = gen_init_iterate + one fixed-point iteration, newton_rel_tol 1e-6.  The reference's committed
    Newton_state.json stops at iteration 2 with Armijo factor 1 for both modules."""
    from nk_ooc_b200.solver import NewtonSolver

    ModelState = _configure(str(tmp_path), "dye_decay_{suff}:001:010")
    assert ModelState.model_config_obj.tracer_module_names == ["dye_decay_001", "dye_decay_010"]
    x = ModelState("gen_init_iterate")
    x += x.comp_fcn(None, None)  # --fp_cnt 1 (test_problem/setup_solver.py:137-155)
    info = dict(TP_SOLVERINFO, newton_rel_tol="1.0e-6")
    solver = NewtonSolver(x, info, workdir=str(tmp_path / "work"), dump=False)
    solver.solve()
    assert solver.iteration == 2
    for rec in solver.history[1:]:
        assert rec["armijo_factor"].shape == (2, 1) and (rec["armijo_factor"] == 1.0).all()
    assert (solver.fcn.norm() < 1.0e-6 * solver.iterate.norm()).all()
    ModelState.reset()


def test_probe_preconditioner_newton_in_one_step(base, tmp_path):
    """the preconditioner built from ONE batched evaluation of coloured column probes
    (solver.ProbePreconditioner; 3 colours x 2 tracers x 20 levels = 120 members): on the grid
    without lateral processes and for the linear iage module M is the exact Jacobian of F, so GMRES
    converges in its first iteration and Newton in one step"""
    from nk_ooc_b200.py_driver_2d.model_state import ModelState
    from nk_ooc_b200.py_driver_2d.setup_solver import gen_grid_vars_file
    from nk_ooc_b200.solver import NewtonSolver, ProbePreconditioner

    info = _modelinfo(str(tmp_path), 20, 3, "0.0", "0.0")
    gen_grid_vars_file(info)
    ModelState.configure(info)
    pre = "ci_py_driver_2d_iage_column_regions/"
    iterate = _state(ModelState, base, pre + "init_iterate")
    made = []

    def factory(it, fcn):
        made.append(ProbePreconditioner(it, fcn))
        return made[-1]

    solver = NewtonSolver(iterate, dict(PD_SOLVERINFO, post_newton_fp_iter="0"), workdir=str(tmp_path / "w0"),
                          dump=False, precond_factory=factory)
    n0 = solver.history[0]["fcn_norm"]
    solver.step()
    rec = solver.history[-1]
    assert made[0].members_probed == 120 and rec["krylov_iterations"] == 1
    # limited by the finite-difference step of the probes and of the Jacobian-vector product
    assert (rec["krylov_precond_resid_norm"][0] < 1e-5 * rec["krylov_beta"]).all()
    assert (rec["fcn_norm"] < 1e-4 * n0).all() and solver.converged_flat()
    # iage is affine in x: the Jacobian of F does not depend on the iterate, one set of probes serves every Newton
    # step (solver.LaggedPrecond) — from a different iterate the same preconditioner converges in one iteration again
    from nk_ooc_b200.solver import LaggedPrecond

    lagged = LaggedPrecond(factory)
    other = _state(ModelState, base, pre + "init_iterate") * 0.5
    solver2 = NewtonSolver(other, dict(PD_SOLVERINFO, post_newton_fp_iter="0", newton_rel_tol="1.0e-12"),
                           workdir=str(tmp_path / "w1"), dump=False, precond_factory=lagged)
    solver2.step()
    solver2.step()
    assert lagged.built == 1 and len(made) == 2
    assert [r["krylov_iterations"] for r in solver2.history[1:]] == [1, 1]
    ModelState.reset()


def test_probe_preconditioned_newton_krylov_with_lateral_processes(tmp_path):
    """Newton-Krylov from gen_init_iterate on the reference's default 40 x 50 grid WITH lateral processes, the
    preconditioner built once from coloured probes that keep the coupling to two columns on each side (5 colours x 2
    tracers x 40 levels = 400 members in one batched evaluation, panel substitution of the block band) and reused for
    every Newton step (the Jacobian of F does not depend on the iterate for iage): converges to newton_rel_tol"""
    from nk_ooc_b200.py_driver_2d.model_state import ModelState
    from nk_ooc_b200.py_driver_2d.setup_solver import gen_grid_vars_file
    from nk_ooc_b200.solver import LaggedPrecond, NewtonSolver, ProbePreconditioner

    info = _modelinfo(str(tmp_path), 40, 50)
    gen_grid_vars_file(info)
    ModelState.configure(info)
    try:
        made = []

        def factory(it, fcn):
            made.append(ProbePreconditioner(it, fcn, reach=2))
            return made[-1]

        lagged = LaggedPrecond(factory)
        solverinfo = dict(PD_SOLVERINFO, krylov_rel_tol="1.0e-3", newton_max_iter="6")
        solver = NewtonSolver(ModelState("gen_init_iterate"), solverinfo, workdir=str(tmp_path / "w"), dump=False,
                              precond_factory=lagged)
        start = float((solver.fcn.norm() / solver.iterate.norm()).max())
        steps = 0
        while not solver.converged_flat() and steps < 4:
            solver.step()
            steps += 1
        rel = float((solver.fcn.norm() / solver.iterate.norm()).max())
        assert solver.converged_flat() and rel < 1.0e-5 < 1.0e-2 < start, (steps, rel)
        assert lagged.built == 1 and made[0].members_probed == 5 * 2 * 40 and made[0]._factor.path == "panel"
        assert all(rec["krylov_iterations"] <= 40 for rec in solver.history[1:])
    finally:
        ModelState.reset()


def test_cli_setup_solver_and_nk_driver_column_regions(base, tmp_path):
    """scripts/ci_py_driver_2d_iage_column_regions.sh through the command line: setup_solver (grid
    file, gen_init_iterate + 1 fixed-point iteration) and nk_driver write the reference's files;
    grid_vars, init_iterate (rtol 1e-3 / atol 1e-6) and iterate_01 (rtol 1.9e-2) against the
    baselines; run_cmd comp_fcn file-to-file"""
    from scipy.io import netcdf_file

    from nk_ooc_b200 import cli
    from nk_ooc_b200.py_driver_2d.model_state import ModelState

    work = str(tmp_path / "work")
    common = ["--model_name", "py_driver_2d", "--workdir", work, "--depth_nlevs", "20", "--ypos_nlevs", "3",
              "--max_abs_vvel", "0.0", "--horiz_mix_coeff", "0.0"]
    assert cli.main(["setup_solver", "--fp_cnt", "1"] + common) == 0
    pre = "ci_py_driver_2d_iage_column_regions/"

    def read(fname, names=("iage", "iage_slow_rest")):
        with netcdf_file(fname, "r", mmap=False) as f:
            return np.stack([np.array(f.variables[n].data) for n in names])

    with netcdf_file(os.path.join(work, "grid_vars.nc"), "r", mmap=False) as f:
        np.testing.assert_array_equal(np.array(f.variables["region_mask"].data), base[pre + "grid_vars/region_mask"])
    np.testing.assert_allclose(read(os.path.join(work, "gen_init_iterate", "init_iterate_0000.nc")),
                               _want(base, pre + "init_iterate_0000"), rtol=1e-7, atol=2e-9)
    np.testing.assert_allclose(read(os.path.join(work, "gen_init_iterate", "init_iterate.nc")),
                               _want(base, pre + "init_iterate"), rtol=1e-3, atol=1e-6)
    ModelState.reset()
    assert cli.main(["nk_driver", "--newton_max_iter", "5"] + common) == 0
    np.testing.assert_allclose(read(os.path.join(work, "iterate_01.nc")), _want(base, pre + "iterate_01"),
                               rtol=1.9e-2, atol=1e-9)
    assert os.path.exists(os.path.join(work, "krylov_00", "krylov_res_00.nc"))
    # --resume on the finished solve: reads the last iterate / fcn back, no further Newton step
    import json

    it_done = json.load(open(os.path.join(work, "Newton_state.json")))["iteration"]
    ModelState.reset()
    assert cli.main(["nk_driver", "--newton_max_iter", "5", "--resume"] + common) == 0
    assert json.load(open(os.path.join(work, "Newton_state.json")))["iteration"] == it_done
    # file-to-file function evaluation
    assert cli.main(["comp_fcn", "--fname_dir", work, "--in_fname", "gen_init_iterate/init_iterate_0000.nc",
                     "--res_fname", "fcn_cli.nc"] + common) == 0
    np.testing.assert_allclose(read(os.path.join(work, "fcn_cli.nc")), _want(base, pre + "fcn_0000"), rtol=1e-3,
                               atol=1e-6)
    ModelState.reset()


def test_newton_state_stats_files_and_resume(base, tmp_path):
    """the work directory of a dumped solve carries the reference's Newton_state.json (iteration 3 and
    the reference's step strings, baselines/ci_long_iage/Newton_state.json) and Newton_stats.nc /
    Krylov_stats.nc; a resumed solver reads iterate and fcn of the logged iteration back instead of
    recomputing them, and a solve interrupted after one step continues to the same answer"""
    import json

    from scipy.io import netcdf_file

    from nk_ooc_b200 import _lib
    from nk_ooc_b200.solver import NewtonSolver

    ModelState = _configure(str(tmp_path), "iage")
    work = str(tmp_path / "work")
    iterate = ModelState({"iage": base["ci_short/init_iterate/iage"]})
    solver = NewtonSolver(iterate, TP_SOLVERINFO, workdir=work)
    solver.step()
    # "interrupted" here: a second solver resumes from the files and finishes
    lib = _lib.load()
    n0 = lib.nkb_launch_count()
    resumed = NewtonSolver(ModelState("zeros"), TP_SOLVERINFO, workdir=work, resume=True)
    assert lib.nkb_launch_count() - n0 <= 4, "resume re-evaluated the function"  # only the two norms of _record
    assert resumed.iteration == 1
    np.testing.assert_array_equal(resumed.iterate.get_tracer_vals("iage"), solver.iterate.get_tracer_vals("iage"))
    np.testing.assert_array_equal(resumed.fcn.get_tracer_vals("iage"), solver.fcn.get_tracer_vals("iage"))
    resumed.solve()
    solver.solve()
    assert resumed.iteration == solver.iteration == 3
    np.testing.assert_allclose(resumed.iterate.get_tracer_vals("iage"), solver.iterate.get_tracer_vals("iage"),
                               rtol=0, atol=1e-12 * np.abs(solver.iterate.get_tracer_vals("iage")).max())
    state = json.load(open(os.path.join(work, "Newton_state.json")))
    assert state["iteration"] == 3
    want_log = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "Newton_state_ci_long_iage.json")))
    ours = set(s.replace(work, "W") for s in state["step_log"])
    for step in want_log["step_log"]:
        step = step.replace("HOME/ci_long_iage_workdir", "W")
        if step.startswith(("__init__", "Newton iterate 0 written")) or ":inc_iteration" in step or \
                (":comp_fcn complete for W/fcn_" in step):
            assert step in ours, step
    with netcdf_file(os.path.join(work, "Newton_stats.nc"), "r", mmap=False) as f:
        assert f.variables["iteration"].shape[0] == 4
        fn = np.array(f.variables["fcn_norm_iage"].data)[:, 0]
        assert (np.diff(fn) < 0).all() and fn[-1] < 1e-8 * np.array(f.variables["iterate_norm_iage"].data)[-1, 0]
        assert (np.array(f.variables["Krylov_iterations"].data)[:3] >= 1).all()
        np.testing.assert_array_equal(np.array(f.variables["Armijo_factor_iage"].data)[:3, 0], 1.0)
        # the model's own statistics: time mean of the tracer over each iteration's hist file
        assert f.variables["iage"].dimensions[0] == "iteration" and f.variables["iage"].shape[0] == 4
        stat = np.array(f.variables["iage"].data)
    with netcdf_file(os.path.join(work, "hist_03.nc"), "r", mmap=False) as f:
        hv = np.array(f.variables["iage"].data)
        wts = np.full(hv.shape[0], 1.0 / (hv.shape[0] - 1))
        wts[0] *= 0.5
        wts[-1] *= 0.5
        np.testing.assert_allclose(stat[3], np.einsum("i,i...", wts, hv), rtol=1e-14)
    with netcdf_file(os.path.join(work, "krylov_00", "Krylov_stats.nc"), "r", mmap=False) as f:
        beta = float(np.array(f.variables["precond_rhs_norm_iage"].data)[0])
        res = np.array(f.variables["precond_resid_norm_iage"].data)[:, 0]
        assert res[-1] < 0.01 * beta
    kstate = json.load(open(os.path.join(work, "krylov_00", "Krylov_state.json")))
    assert "beta" in kstate and "h_mat" in kstate and kstate["iteration"] == len(res)
    ModelState.reset()
