"""The reference's six CI flows (scripts/ci_*.sh of klindsay28/Newton-Krylov_OOC) through their ports in
scripts/ci_*.sh: command line -> GPU model year / Newton-Krylov -> files -> `baseline_cmp` (metadata AND values,
with each script's own tolerances) against the reference's committed baselines, and `diff` of Newton_state.json
(iteration count, the complete step log, Armijo factors) against the baselines' files.

The baselines directory is rebuilt from tests/golden (tests/baseline_files.py); HOME is a scratch directory, as
the scripts put their work directories under $HOME like the reference's.  SURVEY.md §8 f-2 / f-3."""
import os
import subprocess

import pytest

from baseline_files import materialise

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ci_env(tmp_path_factory):
    home = tmp_path_factory.mktemp("home")
    env = dict(os.environ, HOME=str(home), NKB_BASELINES=materialise(str(home / "baselines")))
    return env


def _run(script, env):
    res = subprocess.run(["bash", os.path.join(ROOT, "scripts", script)], env=env, capture_output=True, text=True,
                         timeout=1500)
    logdir = os.environ.get("NKB_CI_LOGDIR")
    if logdir:
        os.makedirs(logdir, exist_ok=True)
        with open(os.path.join(logdir, script + ".log"), "w") as fptr:
            fptr.write(res.stdout + "\n--- stderr ---\n" + res.stderr)
    tail = "\n".join((res.stdout + "\n" + res.stderr).splitlines()[-60:])
    assert res.returncode == 0, f"{script}: err_cnt={res.returncode}\n{tail}"
    assert "err_cnt=0" in res.stdout
    return res.stdout


def test_ci_short(ci_env):
    """depth_axis.nc, fcn_00 / hist_00 / init_iterate / init_iterate_00 at rtol 1e-7, atol 2e-9"""
    _run("ci_short.sh", ci_env)


def test_ci_long_iage(ci_env):
    """needs test_ci_short's work directory (the script compares the two hist files, as the reference's does)"""
    if not os.path.isdir(os.path.join(ci_env["HOME"], "ci_short_workdir")):
        _run("ci_short.sh", ci_env)
    _run("ci_long_iage.sh", ci_env)


def test_ci_long_dye_decay(ci_env):
    _run("ci_long_dye_decay.sh", ci_env)


def test_ci_zero_iage(ci_env):
    _run("ci_zero_iage.sh", ci_env)


def test_ci_py_driver_2d_iage(ci_env):
    _run("ci_py_driver_2d_iage.sh", ci_env)


def test_ci_py_driver_2d_iage_column_regions(ci_env):
    _run("ci_py_driver_2d_iage_column_regions.sh", ci_env)


def test_nk_driver_resumes_and_rewinds_at_step_granularity(ci_env):
    """solver_state.py:36-45,91-98 through the command line: cut Newton_state.json of the finished ci_long_iage
    solve back to the middle of Newton iteration 1 (after its Krylov solve), `nk_driver --resume` finishes from the
    files with the same step log and the same iterate; `--resume --rewind` redoes the last logged step"""
    import json
    import shutil

    import numpy as np
    from scipy.io import netcdf_file

    home = ci_env["HOME"]
    src = os.path.join(home, "ci_long_iage_workdir")
    if not os.path.isdir(src):
        _run("ci_short.sh", ci_env)
        _run("ci_long_iage.sh", ci_env)
    work = os.path.join(home, "resume_workdir")
    shutil.copytree(src, work)
    with open(os.path.join(work, "Newton_state.json")) as fptr:
        text = fptr.read().replace(src, work)
    full = json.loads(text)
    cut = full["step_log"].index("01:_comp_increment complete") + 1
    state = dict(full, iteration=1, step_log=full["step_log"][:cut])
    for key in ("armijo_ind", "armijo_factor", "fp_iter"):
        state.pop(key, None)
    with open(os.path.join(work, "Newton_state.json"), "w") as fptr:
        json.dump(state, fptr, indent=2)
    for fname in ("iterate_02.nc", "iterate_03.nc", "fcn_02.nc", "fcn_03.nc"):
        os.remove(os.path.join(work, fname))
    shutil.rmtree(os.path.join(work, "krylov_02"))
    env = dict(ci_env, PYTHONPATH=os.path.join(ROOT, "newton-krylov_ooc_b200"))
    cmd = ["python", "-m", "nk_ooc_b200.cli", "nk_driver", "--model_name", "test_problem", "--depth_nlevs", "20",
           "--tracer_module_names", "iage", "--workdir", work, "--resume"]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    with open(os.path.join(work, "Newton_state.json")) as fptr:
        resumed = json.load(fptr)
    assert resumed["iteration"] == 3 and resumed["step_log"] == json.loads(text)["step_log"]
    # the Krylov solve of iteration 1 was NOT repeated: its directory still holds the first run's files only
    assert not os.path.exists(os.path.join(work, "krylov_01", "basis_%02d.nc" % 50))

    def iage(fname):
        with netcdf_file(fname, "r", mmap=False) as fptr:
            return np.array(fptr.variables["iage"].data)

    np.testing.assert_allclose(iage(os.path.join(work, "iterate_03.nc")), iage(os.path.join(src, "iterate_03.nc")),
                               rtol=0, atol=1e-11 * np.abs(iage(os.path.join(src, "iterate_03.nc"))).max())
    # rewind: the last logged step ("03:ModelStateBase.put_stats_vars") is redone, the log ends up the same
    res = subprocess.run(cmd + ["--rewind"], env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    with open(os.path.join(work, "Newton_state.json")) as fptr:
        assert json.load(fptr)["step_log"] == resumed["step_log"]


def test_nk_driver_resumes_inside_a_krylov_solve(ci_env):
    """an interruption INSIDE a Krylov solve (krylov_solver.py:105-165 over solver_state.py): Krylov_state.json cut back
    to the end of its first iteration, Newton_state.json to "KrylovSolver instantiated" — the resumed solve reads beta,
    h_mat, the basis vectors and the preconditioned products of the completed iteration back from the files
    (basis_jj.nc, w_jj.nc, precond_fcn_00.nc), continues with iteration 1 and ends with the same increment and the same
    Newton step log"""
    import json
    import shutil

    import numpy as np
    from scipy.io import netcdf_file

    home = ci_env["HOME"]
    src = os.path.join(home, "ci_long_iage_workdir")
    if not os.path.isdir(src):
        _run("ci_short.sh", ci_env)
        _run("ci_long_iage.sh", ci_env)
    # a Newton iteration whose Krylov solve took at least two iterations
    pick = None
    for it in range(3):
        with open(os.path.join(src, f"krylov_{it:02}", "Krylov_state.json")) as fptr:
            if json.load(fptr)["iteration"] >= 2:
                pick = it
                break
    if pick is None:
        pytest.skip("every Krylov solve of this run converged in one iteration")
    work = os.path.join(home, "resume_krylov_workdir")
    shutil.copytree(src, work)
    kdir = os.path.join(work, f"krylov_{pick:02}")
    with open(os.path.join(kdir, "Krylov_state.json")) as fptr:
        ktext = fptr.read().replace(src, work)
    kfull = json.loads(ktext)
    kcut = kfull["step_log"].index("01:inc_iteration") + 1
    h_mat = np.asarray(kfull["h_mat"]["__ndarray__"])
    kstate = {"iteration": 1, "step_log": kfull["step_log"][:kcut], "beta": kfull["beta"],
              "h_mat": {"__ndarray__": h_mat[:, :2, :1, :].tolist()}}
    with open(os.path.join(kdir, "Krylov_state.json"), "w") as fptr:
        json.dump(kstate, fptr, indent=2)
    for j in range(1, kfull["iteration"]):  # files of the iterations that "did not happen"
        for quantity in ("w_raw", "w", "krylov_res", "perturb_fcn_w_raw"):
            fname = os.path.join(kdir, f"{quantity}_{j:02}.nc")
            if os.path.exists(fname):
                os.remove(fname)
        later = os.path.join(kdir, f"basis_{j + 1:02}.nc")
        if os.path.exists(later):
            os.remove(later)
    with open(os.path.join(work, "Newton_state.json")) as fptr:
        ntext = fptr.read().replace(src, work)
    nfull = json.loads(ntext)
    ncut = nfull["step_log"].index(f"{pick:02}:KrylovSolver instantiated") + 1
    nstate = {"iteration": pick, "step_log": nfull["step_log"][:ncut]}
    with open(os.path.join(work, "Newton_state.json"), "w") as fptr:
        json.dump(nstate, fptr, indent=2)
    for later in range(pick + 1, 4):
        for fname in (f"iterate_{later:02}.nc", f"fcn_{later:02}.nc", f"hist_{later:02}.nc"):
            if os.path.exists(os.path.join(work, fname)):
                os.remove(os.path.join(work, fname))
        shutil.rmtree(os.path.join(work, f"krylov_{later:02}"), ignore_errors=True)
    os.remove(os.path.join(work, f"increment_{pick:02}.nc"))
    env = dict(ci_env, PYTHONPATH=os.path.join(ROOT, "newton-krylov_ooc_b200"))
    cmd = ["python", "-m", "nk_ooc_b200.cli", "nk_driver", "--model_name", "test_problem", "--depth_nlevs", "20",
           "--tracer_module_names", "iage", "--workdir", work, "--resume"]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    with open(os.path.join(work, "Newton_state.json")) as fptr:
        resumed = json.load(fptr)
    assert resumed["iteration"] == nfull["iteration"] and resumed["step_log"] == nfull["step_log"]
    with open(os.path.join(kdir, "Krylov_state.json")) as fptr:
        kres = json.load(fptr)
    assert kres["iteration"] == kfull["iteration"] and kres["step_log"] == kfull["step_log"]
    np.testing.assert_allclose(np.asarray(kres["h_mat"]["__ndarray__"]), h_mat, rtol=1e-9, atol=1e-12 * np.abs(h_mat).max())

    def iage(fname):
        with netcdf_file(fname, "r", mmap=False) as fptr:
            return np.array(fptr.variables["iage"].data)

    for fname in (f"increment_{pick:02}.nc", f"iterate_{nfull['iteration']:02}.nc"):
        want = iage(os.path.join(src, fname))
        np.testing.assert_allclose(iage(os.path.join(work, fname)), want, rtol=0, atol=1e-9 * np.abs(want).max())
