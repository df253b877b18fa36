"""CPU tests of the host logic of the Newton-Krylov driver (no device needed)"""
import numpy as np


def test_krylov_basis_coeffs_least_squares():
    """krylov_solver.py:168-182: per (module, region) min || beta e1 - H y ||"""
    from nk_ooc_b200.solver import comp_krylov_basis_coeffs

    rng = np.random.default_rng(0)
    n_mod, j, R = 2, 3, 4
    h_mat = np.zeros((n_mod, j + 2, j + 1, R))
    for m in range(n_mod):
        for r in range(R):
            h = np.triu(rng.normal(size=(j + 2, j + 1)), -1)  # upper Hessenberg
            h_mat[m, :, :, r] = h
    beta = np.abs(rng.normal(size=(n_mod, R))) + 0.1
    coeff = comp_krylov_basis_coeffs(beta, h_mat)
    assert coeff.shape == (n_mod, j + 1, R)
    for m in range(n_mod):
        for r in range(R):
            h = h_mat[m, :, :, r]
            rhs = np.zeros(j + 2)
            rhs[0] = beta[m, r]
            want = np.linalg.solve(h.T @ h, h.T @ rhs)
            np.testing.assert_allclose(coeff[m, :, r], want, rtol=1e-9)
    # one iteration: y = beta h00 / (h00^2 + h10^2)
    h1 = np.zeros((1, 2, 1, 1))
    h1[0, :, 0, 0] = [3.0, 4.0]
    np.testing.assert_allclose(comp_krylov_basis_coeffs(np.array([[5.0]]), h1)[0, 0, 0], 5.0 * 3.0 / 25.0)


# ---- the Newton / Krylov control flow over a numpy stand-in for the model state (no device) ---------------
import json
import os
from types import SimpleNamespace

import pytest


from fake_state import FakeState as _FakeState, Interrupted as _Interrupted  # noqa: E402


SOLVERINFO = {"newton_rel_tol": "1.0e-8", "newton_max_iter": "12", "post_newton_fp_iter": "1", "krylov_rel_tol": "0.01"}


@pytest.fixture
def fake(monkeypatch):
    from nk_ooc_b200 import model_state_base

    def lin_comb(cls, coeff, fname_fcn, quantity):
        res = cls(np.zeros(6))
        for i in range(coeff.shape[1]):
            res.vals += coeff[:, i, :, None] * fname_fcn(quantity, i).vals
        return res

    monkeypatch.setattr(model_state_base, "lin_comb", lin_comb)
    _FakeState.configure("mild")
    yield _FakeState
    _FakeState.configure("mild")


def _solve(cls, workdir, **kw):
    from nk_ooc_b200.solver import NewtonSolver

    # the damped problem runs without the fixed-point iterations (x + F(x) is no contraction there): this is also the
    # path on which an accepted Armijo candidate's F becomes the next iteration's F without another evaluation
    info = dict(SOLVERINFO, post_newton_fp_iter="0") if cls.problem == "damped" else SOLVERINFO
    solver = NewtonSolver(cls(np.ones(6)), info, workdir=workdir, **kw)
    solver.solve()
    return solver


def test_newton_krylov_control_flow_and_reference_step_log(fake, tmp_path):
    """the solvers converge on the stand-in problem, a solve that keeps no files gives the same iterate, and the
    step log of a kept solve is, iteration by iteration, the sequence of the reference's own run
    (baselines/ci_long_iage/Newton_state.json)"""
    work = str(tmp_path / "w")
    solver = _solve(fake, work)
    assert solver.converged_flat() and 3 <= solver.iteration <= 10
    x = solver.iterate.vals
    np.testing.assert_allclose(fake.A @ x[0, 0] - fake.b + 0.02 * x[0, 0] ** 3, 0.0, atol=1e-6)
    nofiles = _solve(fake, str(tmp_path / "n"), dump=False)
    np.testing.assert_allclose(nofiles.iterate.vals, x, rtol=0, atol=1e-12)
    assert not os.path.exists(str(tmp_path / "n" / "Newton_state.json"))
    with open(os.path.join(work, "Newton_state.json")) as fptr:
        state = json.load(fptr)
    golden = os.path.join(os.path.dirname(__file__), "golden", "Newton_state_ci_long_iage.json")
    with open(golden) as fptr:
        want = json.load(fptr)["step_log"]
    ours = [s.replace(work, "HOME/ci_long_iage_workdir") for s in state["step_log"]]
    # the reference's run took 3 Newton iterations; compare the steps of every iteration both runs have
    cut = lambda log: [s for s in log if not s[:2].isdigit() or int(s[:2]) < 3]  # noqa: E731
    assert cut(ours) == cut(want)
    assert state["fp_iter"] == 1 and state["armijo_ind"] == 0


@pytest.mark.parametrize("problem", ["mild", "damped", "regions"])
def test_resume_after_an_interruption_at_every_function_evaluation(fake, tmp_path, problem):
    """solver_state.py:36-45 / newton_solver.py:140-334 / krylov_solver.py:86-165: a solve interrupted at ANY of its
    function evaluations and resumed from the files ends with the iterate and the step log of the uninterrupted solve,
    and only the interrupted evaluation is done twice"""
    steep = problem != "mild"
    fake.configure(problem)
    ref = _solve(fake, str(tmp_path / "ref"))
    total = fake.calls
    assert ref.converged_flat()
    if steep:
        # (the damped steps of this run: Armijo candidates 01.. were evaluated and logged)
        from scipy.io import netcdf_file

        with netcdf_file(str(tmp_path / "ref" / "Newton_stats.nc"), "r", mmap=False) as nc:
            factors = np.array(nc.variables["Armijo_factor_iage"].data)[: ref.iteration]
        assert factors.min() < 1.0 and any("prov_fcn_Armijo_01_" in f for f in os.listdir(str(tmp_path / "ref")))
    with open(str(tmp_path / "ref" / "Newton_state.json")) as fptr:
        ref_log = [s.replace(str(tmp_path / "ref"), "W") for s in json.load(fptr)["step_log"]]
    assert total >= 12
    for k in range(1, total + 1):
        work = str(tmp_path / f"w{k}")
        fake.calls, fake.fail_at = 0, k
        with pytest.raises(_Interrupted):
            _solve(fake, work)
        fake.fail_at = None
        resumed = _solve(fake, work, resume=True)
        assert fake.calls == total + 1, k  # k - 1 before the interruption, the interrupted one, the rest once
        np.testing.assert_allclose(resumed.iterate.vals, ref.iterate.vals, rtol=0, atol=1e-12, err_msg=str(k))
        with open(os.path.join(work, "Newton_state.json")) as fptr:
            assert [s.replace(work, "W") for s in json.load(fptr)["step_log"]] == ref_log, k


def test_rewind_redoes_the_last_logged_step(fake, tmp_path):
    work = str(tmp_path / "w")
    ref = _solve(fake, work)
    with open(os.path.join(work, "Newton_state.json")) as fptr:
        log = json.load(fptr)["step_log"]
    from nk_ooc_b200.solver import NewtonSolver

    fake.calls = 0
    again = NewtonSolver(fake(np.ones(6)), SOLVERINFO, workdir=work, resume=True, rewind=True)
    again.solve()
    assert fake.calls == 0  # the last logged step is the stats put of the final iteration: no evaluation
    np.testing.assert_array_equal(again.iterate.vals, ref.iterate.vals)
    with open(os.path.join(work, "Newton_state.json")) as fptr:
        assert json.load(fptr)["step_log"] == log
    with pytest.raises(RuntimeError):
        NewtonSolver(fake(np.ones(6)), SOLVERINFO, workdir=str(tmp_path / "x"), resume=False, rewind=True)


def _same_stats(path, want):
    from scipy.io import netcdf_file

    with netcdf_file(path, "r", mmap=False) as nc:
        assert hasattr(nc, "history")
        for name, size, length in want["dimensions"]:
            assert name in nc.dimensions and nc.dimensions[name] == size, (path, name)
        assert sorted(nc.variables) == sorted(v["name"] for v in want["variables"]), path
        for var in want["variables"]:
            got = nc.variables[var["name"]]
            assert list(got.dimensions) == var["dimensions"], var["name"]
            attrs = {k: (v.decode() if isinstance(v, bytes) else v) for k, v in got._attributes.items()}  # noqa: SLF001
            assert sorted(attrs) == sorted(var["attrs"]), var["name"]
            for key, val in var["attrs"].items():
                if isinstance(val, str):
                    assert attrs[key] == val, (var["name"], key)
                else:
                    assert float(np.asarray(attrs[key]).reshape(-1)[0]) == float(val), (var["name"], key)
            data = np.array(got.data, dtype=float)
            assert data.shape == np.shape(var["data"]), var["name"]
            np.testing.assert_allclose(data, var["data"], rtol=1e-9, atol=1e-14, err_msg=var["name"])


@pytest.mark.parametrize("problem", ["mild", "damped", "regions", "min_iter"])
def test_same_solve_as_the_references_own_solvers(fake, tmp_path, problem):
    """tests/golden/ref_solver_<problem>.json records what the REFERENCE's NewtonSolver / KrylovSolver (imported
    unmodified, oracle/gen_golden_solver.py) did over this very state class: this package's solvers take the same
    Newton iterates, function values and increments, evaluate F as often, and leave the same Newton and Krylov step
    logs, saved Hessenberg matrices and files behind"""
    with open(os.path.join(os.path.dirname(__file__), "golden", f"ref_solver_{problem}.json")) as fptr:
        ref = json.load(fptr)
    from nk_ooc_b200.solver import NewtonSolver

    fake.configure(ref["problem"])
    work = str(tmp_path / "w")
    solver = NewtonSolver(fake(np.ones(6)), dict(ref["solverinfo"]), workdir=work)
    solver.solve()
    assert solver.iteration == ref["iterations"] and fake.calls == ref["evaluations"]

    def arr(name):
        return fake(os.path.join(work, name)).vals

    for i in range(ref["iterations"] + 1):
        np.testing.assert_allclose(arr(f"iterate_{i:02}.nc"), ref["iterate"][i], rtol=1e-11, atol=1e-13, err_msg=f"iterate {i}")
        np.testing.assert_allclose(arr(f"fcn_{i:02}.nc"), ref["fcn"][i], rtol=1e-9, atol=1e-13, err_msg=f"fcn {i}")
    for i in range(ref["iterations"]):
        np.testing.assert_allclose(arr(f"increment_{i:02}.nc"), ref["increment"][i], rtol=1e-9, atol=1e-13)

    def state(path):
        with open(path) as fptr:
            rec = json.load(fptr)
        rec["step_log"] = [s.replace(work, "W") for s in rec["step_log"]]
        return rec

    def vals(rec, key):
        val = rec[key]
        return np.array(val["__ndarray__"] if isinstance(val, dict) else val, dtype=float)

    def same_schema(got, want, what):
        """the same keys, and arrays stored with the same `__ndarray__` tagging (solver_state.py:14-33,137-166)"""
        assert list(got) == list(want), what
        for key, val in want.items():
            assert isinstance(got[key], type(val)), (what, key)
            if isinstance(val, dict):
                assert list(got[key]) == list(val) == ["__ndarray__"], (what, key)
                assert np.shape(got[key]["__ndarray__"]) == np.shape(val["__ndarray__"]), (what, key)

    ours = state(os.path.join(work, "Newton_state.json"))
    same_schema(ours, ref["Newton_state"], "Newton_state.json")
    assert ours["step_log"] == ref["Newton_state"]["step_log"]
    assert ours["iteration"] == ref["Newton_state"]["iteration"]
    for key in ("armijo_ind", "armijo_factor", "fp_iter"):
        np.testing.assert_array_equal(vals(ours, key), vals(ref["Newton_state"], key), err_msg=key)
    for i, want in enumerate(ref["Krylov_state"]):
        got = state(os.path.join(work, f"krylov_{i:02}", "Krylov_state.json"))
        same_schema(got, want, f"krylov_{i:02}/Krylov_state.json")
        assert got["step_log"] == want["step_log"], f"Krylov solve {i}"
        assert got["iteration"] == want["iteration"]
        np.testing.assert_allclose(vals(got, "beta"), vals(want, "beta"), rtol=1e-10)
        np.testing.assert_allclose(vals(got, "h_mat"), vals(want, "h_mat"), rtol=1e-8, atol=1e-12)
    # the stats files: every dimension, variable, attribute and value the reference wrote (stats_file.py,
    # solver_base.py:68-193, newton_solver.py:62-118, krylov_solver.py:50-73), fill values of the grown iteration
    # dimension included
    same_stats = _same_stats
    same_stats(os.path.join(work, "Newton_stats.nc"), ref["Newton_stats"])
    for i, want in enumerate(ref["Krylov_stats"]):
        same_stats(os.path.join(work, f"krylov_{i:02}", "Krylov_stats.nc"), want)
    # the files left in the work directory (the reference's stats files were kept in memory by the generator)
    files = sorted(os.path.relpath(os.path.join(d, f), work) for d, _, fs in os.walk(work) for f in fs)
    stats = {f for f in files if f.endswith("_stats.nc")}
    assert stats == {"Newton_stats.nc"} | {os.path.join(f"krylov_{i:02}", "Krylov_stats.nc") for i in range(ref["iterations"])}
    assert [f for f in files if f not in stats] == [f for f in ref["files"] if f != "init_iterate.nc"]
    if ref["problem"] != "mild":
        assert 0.25 in np.ravel(ref["Armijo_factor"])  # a damped step ...
    if ref["problem"] == "regions":
        assert 0.0 in np.ravel(ref["Armijo_factor"])  # ... and blocks that had converged while others had not


@pytest.mark.parametrize("problem", ["mild", "damped", "regions", "min_iter"])
def test_the_reference_resumes_a_solve_this_package_interrupted(fake, tmp_path, problem):
    """state-file compatibility in the direction a user switching back would need: a solve of THIS package's solvers,
    interrupted at a function evaluation, is picked up by the REFERENCE's `NewtonSolver(resume=True)` (build container
    only: skipped where /root/reference is absent) — Newton_state.json, the Krylov_state.json of a solve in progress and
    the iterate / basis / w files are read by the reference's own code, which finishes with the iterate and the Newton
    step log of its own uninterrupted solve, evaluating only what was not logged"""
    from oracle import gen_golden_solver as gen
    from oracle import ref_harness

    if not ref_harness.available():
        pytest.skip("the reference is not mounted here")
    with open(os.path.join(os.path.dirname(__file__), "golden", f"ref_solver_{problem}.json")) as fptr:
        ref = json.load(fptr)
    from nk_ooc_b200.solver import NewtonSolver

    ref_solver_class = gen.reference_newton_solver()
    for k in (2, 5, 9, 14, ref["evaluations"] - 1):
        fake.configure(ref["problem"])
        work = str(tmp_path / f"w{k}")
        fake.calls, fake.fail_at = 0, k
        with pytest.raises(_Interrupted):
            NewtonSolver(fake(np.ones(6)), dict(ref["solverinfo"]), workdir=work).solve()
        fake.fail_at = None
        theirs = ref_solver_class(fake, gen.solverinfo(work, **gen.CASES[problem][1]), resume=True, rewind=False)
        while not theirs.converged().all():
            theirs.step()
        assert fake.calls == ref["evaluations"] + 1, k
        final = fake(os.path.join(work, f"iterate_{ref['iterations']:02}.nc")).vals
        np.testing.assert_allclose(final, ref["iterate"][-1], rtol=1e-11, atol=1e-13, err_msg=str(k))
        with open(os.path.join(work, "Newton_state.json")) as fptr:
            log = [s.replace(work, "W") for s in json.load(fptr)["step_log"]]
        assert log == ref["Newton_state"]["step_log"], k


@pytest.mark.parametrize("problem", ["mild", "damped", "regions", "min_iter"])
def test_this_package_resumes_a_solve_the_reference_was_interrupted_in(fake, tmp_path, problem, monkeypatch):
    """the direction a user switching over needs (build container only): the REFERENCE's NewtonSolver is interrupted at
    a function evaluation; `NewtonSolver(resume=True)` of this package reads the reference's Newton_state.json,
    Krylov_state.json, stats files, iterate / basis / w files and finishes with the iterate, the step log and the stats
    files of the reference's own uninterrupted solve, evaluating only what was not logged"""
    from oracle import gen_golden_solver as gen
    from oracle import ref_harness

    if not ref_harness.available():
        pytest.skip("the reference is not mounted here")
    with open(os.path.join(os.path.dirname(__file__), "golden", f"ref_solver_{problem}.json")) as fptr:
        ref = json.load(fptr)
    from nk_ooc_b200.solver import NewtonSolver

    monkeypatch.setattr(gen, "PERSIST", True)  # the reference's stats files as real files
    ref_solver_class = gen.reference_newton_solver()
    for k in (2, 5, 9, 14, ref["evaluations"] - 1):
        fake.configure(ref["problem"])
        work = str(tmp_path / f"w{k}")
        init = os.path.join(work, "init_iterate.nc")
        fake(np.ones(6)).dump(init)
        fake.fail_at = k
        with pytest.raises(_Interrupted):
            theirs = ref_solver_class(fake, gen.solverinfo(work, init, **gen.CASES[problem][1]), resume=False, rewind=False)
            while not theirs.converged().all():
                theirs.step()
        fake.fail_at = None
        ours = NewtonSolver(fake(np.ones(6)), dict(ref["solverinfo"]), workdir=work, resume=True)
        ours.solve()
        assert fake.calls == ref["evaluations"] + 1, k
        assert ours.iteration == ref["iterations"]
        np.testing.assert_allclose(ours.iterate.vals, ref["iterate"][-1], rtol=1e-11, atol=1e-13, err_msg=str(k))
        with open(os.path.join(work, "Newton_state.json")) as fptr:
            log = [s.replace(work, "W") for s in json.load(fptr)["step_log"]]
        assert log == ref["Newton_state"]["step_log"], k
        _same_stats(os.path.join(work, "Newton_stats.nc"), ref["Newton_stats"])
        for i, want in enumerate(ref["Krylov_stats"]):
            _same_stats(os.path.join(work, f"krylov_{i:02}", "Krylov_stats.nc"), want)


@pytest.mark.parametrize("problem", ["damped", "regions"])
def test_rewind_after_an_interruption_in_both_implementations(fake, tmp_path, problem, monkeypatch):
    """`--resume --rewind` (solver_state.py:36-45,91-98; newton_solver.py:158-166: a rewound "KrylovSolver instantiated"
    step rewinds the Krylov solver too): the last logged step of an interrupted solve is taken back and redone.  This
    package's solver ends with the iterate of the uninterrupted solve; in the build container the SAME interrupted
    work directory (written by either implementation) is also handed to the reference's solver, and both leave the
    same Newton step log behind (a popped "inc_iteration" is not logged again by either: the counter was saved)."""
    import shutil

    from oracle import gen_golden_solver as gen
    from oracle import ref_harness

    with open(os.path.join(os.path.dirname(__file__), "golden", f"ref_solver_{problem}.json")) as fptr:
        ref = json.load(fptr)
    from nk_ooc_b200.solver import NewtonSolver

    have_ref = ref_harness.available()
    if have_ref:
        monkeypatch.setattr(gen, "PERSIST", True)
        ref_solver_class = gen.reference_newton_solver()

    def ours(work, **kw):
        NewtonSolver(fake(np.ones(6)), dict(ref["solverinfo"]), workdir=work, **kw).solve()

    def theirs(work, init, **kw):
        solver = ref_solver_class(fake, gen.solverinfo(work, init), **kw)
        while not solver.converged().all():
            solver.step()

    def outcome(work, what):
        final = fake(os.path.join(work, f"iterate_{ref['iterations']:02}.nc")).vals
        np.testing.assert_allclose(final, ref["iterate"][-1], rtol=1e-11, atol=1e-13, err_msg=what)
        with open(os.path.join(work, "Newton_state.json")) as fptr:
            return [s.replace(work, "W") for s in json.load(fptr)["step_log"]]

    for k in range(2, ref["evaluations"], 3):
        for first in ["ours"] + (["theirs"] if have_ref else []):
            fake.configure(ref["problem"])
            work = str(tmp_path / f"{first}_{k}")
            init = os.path.join(work, "init_iterate.nc")
            fake(np.ones(6)).dump(init)
            fake.fail_at = k
            with pytest.raises(_Interrupted):
                if first == "ours":
                    ours(work)
                else:
                    theirs(work, init, resume=False, rewind=False)
            fake.fail_at = None
            what = f"{first} interrupted at evaluation {k}"
            if have_ref:
                # the reference first, in place (the step strings hold the directory's path), then the directory is put
                # back as the interruption left it
                shutil.copytree(work, work + "_bak")
                theirs(work, None, resume=True, rewind=True)
                want = outcome(work, what + ", the reference resumed with rewind")
                shutil.rmtree(work)
                shutil.copytree(work + "_bak", work)
            ours(work, resume=True, rewind=True)
            got = outcome(work, what + ", this package resumed with rewind")
            if have_ref:
                assert got == want, what
