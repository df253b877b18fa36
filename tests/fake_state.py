"""A numpy stand-in for a model state with the operator surface the Newton / Krylov solvers use — shared by
tests/test_solver_host.py (this package's solvers) and oracle/gen_golden_solver.py (the REFERENCE's own solvers, run in
the build container over this very class to produce tests/golden/ref_solver_*.json)."""
import os
from types import SimpleNamespace

import numpy as np


class Interrupted(Exception):
    pass


class FakeState:
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
            # (negative and small in slope, like F = x(T) - x(0) of a dissipative model: the post-Newton fixed-point
            # iteration x + F(x) is a contraction)
            resid = cls.A @ x - cls.b
            return -(0.05 * np.arctan(3.0 * resid) + 0.005 * resid)
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
        return FakeState(self.vals.copy() if clone_vals else np.zeros_like(self.vals))

    # the model
    def comp_fcn(self, res_fname, solver_state, hist_fname=None):
        step = f"comp_fcn complete for {res_fname}"
        if solver_state is not None and solver_state.step_logged(step):
            return FakeState(res_fname)
        cls = type(self)
        cls.calls += 1
        if cls.fail_at is not None and cls.calls == cls.fail_at:
            raise Interrupted(step)
        res = FakeState(type(self).fcn_of(self.vals))
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
            return FakeState(res_fname)
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
        # (model_state_base.py:404-406: the reference's method logs this step itself; this package's solver logs it)
        step = f"ModelStateBase.gen_precond_jacobian {precond_fname}"
        if solver_state is not None:
            if solver_state.step_logged(step, per_iteration=False):
                return
            solver_state.log_step(step, per_iteration=False)
        assert os.path.exists(hist_fname)
        os.makedirs(os.path.dirname(precond_fname), exist_ok=True)
        with open(precond_fname, "w") as fptr:
            fptr.write("precond")

    def apply_precond_jacobian(self, precond_fname, res_fname, solver_state):
        step = f"apply_precond_jacobian complete for {res_fname}"
        if solver_state is not None and solver_state.step_logged(step):
            return FakeState(res_fname)
        res = FakeState(self.vals / np.diag(self.A)).dump(res_fname, "precond")
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
            if not isinstance(v, FakeState):  # (the reference hands file names, this package's solver resident states)
                v = FakeState(v)
            h[:, i, :] = self.dot_prod(v)
            self.vals -= h[0, i, 0] * v.vals
        return h

    # operators with [n_modules, region_cnt] scalars
    @staticmethod
    def _s(other):
        return float(np.asarray(other).reshape(-1)[0]) if not isinstance(other, FakeState) else other.vals

    def __neg__(self):
        return FakeState(-self.vals)

    def __add__(self, other):
        return FakeState(self.vals + self._s(other))

    def __sub__(self, other):
        return FakeState(self.vals - self._s(other))

    def __mul__(self, other):
        return FakeState(self.vals * self._s(other))

    __rmul__ = __mul__

    def __truediv__(self, other):
        return FakeState(self.vals / self._s(other))

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

    def log(self, msg=None):
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
