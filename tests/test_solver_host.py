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


class _Interrupted(Exception):
    pass


class _FakeState:
    """the operator surface solver.py uses, on a 6-vector: F(x) = A x - b + 0.02 x^3, preconditioner diag(A)^-1,
    files are .npy arrays under the reference's file names, steps are logged exactly where the model states of
    this package (and the reference's) log them"""

    __array_priority__ = 100
    model_config_obj = SimpleNamespace(region_cnt=1)
    rng = np.random.default_rng(7)
    A = np.diag(np.linspace(2.0, 5.0, 6)) + 0.3 * rng.normal(size=(6, 6))
    b = rng.normal(size=6)
    calls = 0
    fail_at = None
    members = 1
    steep = False

    @classmethod
    def fcn_of(cls, x):
        if cls.steep:
            # Newton's full step overshoots from x = 1 (the arctangent flattens): the Armijo loop has to damp it
            return np.arctan(3.0 * (cls.A @ x - cls.b)) + 0.1 * (cls.A @ x - cls.b)
        return cls.A @ x - cls.b + 0.02 * x ** 3

    def __init__(self, vals):
        if isinstance(vals, str):
            with open(vals, "rb") as fptr:
                vals = np.load(fptr)
        self.vals = np.array(vals, dtype=float)
        self.tracer_modules = [SimpleNamespace(name="iage", units="years")]

    # files
    def dump(self, fname, caller=None):
        if fname is not None:
            os.makedirs(os.path.dirname(fname), exist_ok=True)
            with open(fname, "wb") as fptr:
                np.save(fptr, self.vals)
        return self

    def _like(self, clone_vals=True):
        return _FakeState(self.vals.copy() if clone_vals else np.zeros_like(self.vals))

    # the model
    def comp_fcn(self, res_fname, solver_state, hist_fname=None):
        step = f"comp_fcn complete for {res_fname}"
        if solver_state is not None and solver_state.step_logged(step):
            return _FakeState(res_fname)
        cls = type(self)
        cls.calls += 1
        if cls.fail_at is not None and cls.calls == cls.fail_at:
            raise _Interrupted(step)
        res = _FakeState(type(self).fcn_of(self.vals))
        if hist_fname is not None:
            os.makedirs(os.path.dirname(hist_fname), exist_ok=True)
            with open(hist_fname, "w") as fptr:
                fptr.write("hist")
        res.dump(res_fname, "comp_fcn")
        if solver_state is not None:
            solver_state.log_step(step)
        return res

    def comp_jacobian_fcn_state_prod(self, fcn, direction, res_fname, solver_state):
        step = f"comp_jacobian_fcn_state_prod complete for {res_fname}"
        if solver_state is not None and solver_state.step_logged(step):
            return _FakeState(res_fname)
        sigma = 1.0e-4 * self.norm()
        sigma = np.where(sigma == 0.0, 1.0, sigma)
        perturb = self + sigma * direction
        pname = None
        if res_fname is not None:
            pname = os.path.join(os.path.dirname(res_fname), f"perturb_fcn_{os.path.basename(res_fname)}")
        res = ((perturb.comp_fcn(pname, solver_state) - fcn) / sigma).dump(res_fname, "jvp")
        if solver_state is not None:
            solver_state.log_step(step)
        return res

    def gen_precond_jacobian(self, hist_fname, precond_fname, solver_state=None):
        assert os.path.exists(hist_fname)
        os.makedirs(os.path.dirname(precond_fname), exist_ok=True)
        with open(precond_fname, "w") as fptr:
            fptr.write("precond")

    def apply_precond_jacobian(self, precond_fname, res_fname, solver_state):
        step = f"apply_precond_jacobian complete for {res_fname}"
        if solver_state is not None and solver_state.step_logged(step):
            return _FakeState(res_fname)
        res = _FakeState(self.vals / np.diag(self.A)).dump(res_fname, "precond")
        if solver_state is not None:
            solver_state.log_step(step)
        return res

    # reductions
    def dot_prod(self, other):
        return np.array([[np.mean(self.vals * other.vals)]])

    def norm(self):
        return np.sqrt(self.dot_prod(self))

    def mean(self):
        return np.array([[np.mean(self.vals)]])

    def mod_gram_schmidt(self, basis_cnt, fname_fcn, quantity):
        h = np.zeros((1, basis_cnt, 1))
        for i in range(basis_cnt):
            v = fname_fcn(quantity, i)
            h[:, i, :] = self.dot_prod(v)
            self.vals -= h[0, i, 0] * v.vals
        return h

    # operators with [n_modules, region_cnt] scalars
    @staticmethod
    def _s(other):
        return float(np.asarray(other).reshape(-1)[0]) if not isinstance(other, _FakeState) else other.vals

    def __neg__(self):
        return _FakeState(-self.vals)

    def __add__(self, other):
        return _FakeState(self.vals + self._s(other))

    def __sub__(self, other):
        return _FakeState(self.vals - self._s(other))

    def __mul__(self, other):
        return _FakeState(self.vals * self._s(other))

    __rmul__ = __mul__

    def __truediv__(self, other):
        return _FakeState(self.vals / self._s(other))

    def __iadd__(self, other):
        self.vals = self.vals + self._s(other)
        return self

    def __itruediv__(self, other):
        self.vals = self.vals / self._s(other)
        return self

    # the rest of the surface
    def apply_limiter(self, base):
        return np.ones((1, 1))

    def log_vals(self, msg, vals):
        pass

    def copy_real_tracers_to_shadow_tracers(self):
        return self

    def copy_shadow_tracers_to_real_tracers(self):
        return self

    def shadow_tracers_on(self):
        return False

    def _log_only(self, step, solver_state, per_iteration):
        if solver_state is not None:
            solver_state.log_step(step, per_iteration)

    def def_stats_vars(self, stats_file, hist_fname, solver_state):
        self._log_only("ModelStateBase.def_stats_vars", solver_state, False)

    def put_stats_vars_iteration_invariant(self, stats_file, hist_fname, solver_state):
        self._log_only("ModelStateBase.put_stats_vars_iteration_invariant", solver_state, False)

    def put_stats_vars(self, stats_file, hist_fname, solver_state):
        self._log_only("ModelStateBase.put_stats_vars", solver_state, True)


SOLVERINFO = {"newton_rel_tol": "1.0e-8", "newton_max_iter": "12", "post_newton_fp_iter": "1", "krylov_rel_tol": "0.01"}


@pytest.fixture
def fake(monkeypatch):
    from nk_ooc_b200 import model_state_base

    def lin_comb(cls, coeff, fname_fcn, quantity):
        res = cls(np.zeros(6))
        for i in range(coeff.shape[1]):
            res.vals += coeff[0, i, 0] * fname_fcn(quantity, i).vals
        return res

    monkeypatch.setattr(model_state_base, "lin_comb", lin_comb)
    _FakeState.calls, _FakeState.fail_at, _FakeState.steep = 0, None, False
    yield _FakeState
    _FakeState.steep = False


def _solve(cls, workdir, **kw):
    from nk_ooc_b200.solver import NewtonSolver

    # the damped problem runs without the fixed-point iterations (x + F(x) is no contraction there): this is also the
    # path on which an accepted Armijo candidate's F becomes the next iteration's F without another evaluation
    info = dict(SOLVERINFO, post_newton_fp_iter="0") if cls.steep else SOLVERINFO
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
    np.testing.assert_allclose(fake.A @ x - fake.b + 0.02 * x ** 3, 0.0, atol=1e-6)
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


@pytest.mark.parametrize("steep", [False, True])
def test_resume_after_an_interruption_at_every_function_evaluation(fake, tmp_path, steep):
    """solver_state.py:36-45 / newton_solver.py:140-334 / krylov_solver.py:86-165: a solve interrupted at ANY of its
    function evaluations and resumed from the files ends with the iterate and the step log of the uninterrupted solve,
    and only the interrupted evaluation is done twice"""
    fake.steep = steep
    ref = _solve(fake, str(tmp_path / "ref"))
    total = fake.calls
    assert ref.converged_flat()
    if steep:
        # (the damped steps of this run: Armijo candidates 01.. were evaluated and logged)
        from scipy.io import netcdf_file

        with netcdf_file(str(tmp_path / "ref" / "Newton_stats.nc"), "r", mmap=False) as nc:
            factors = np.array(nc.variables["Armijo_factor_iage"].data)[: ref.iteration, 0]
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
