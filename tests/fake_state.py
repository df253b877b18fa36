"""A numpy stand-in for a model state with the operator surface the Newton / Krylov solvers use — shared by
tests/test_solver_host.py (this package's solvers) and oracle/gen_golden_solver.py (the REFERENCE's own solvers, run in
the build container over this very class to produce tests/golden/ref_solver_*.json).

A state is an array [n_modules, region_cnt, 6]: every (tracer module, region) pair is an independent 6-vector problem
F(x) = 0 of its own kind, so that norms, Armijo factors, Krylov coefficients and convergence are per module and region
as in the reference (arrays [n_modules, region_cnt] throughout).  Files are .npy arrays under the reference's file
names; steps are logged exactly where the model states of this package (and the reference's) log them."""
import os
from types import SimpleNamespace

import numpy as np

NVEC = 6
_RNG = np.random.default_rng(7)
_A0 = np.diag(np.linspace(2.0, 5.0, NVEC)) + 0.3 * _RNG.normal(size=(NVEC, NVEC))
_B0 = _RNG.normal(size=NVEC)
# [module][region] -> kind of the block's function
PROBLEMS = {
    "mild": [["cubic"]],
    "damped": [["arctan"]],
    # two tracer modules x two regions: one needs Armijo damping, one is linear (converged after the first step, its
    # Armijo factor then 0), the others are mildly nonlinear with different matrices
    "regions": [["cubic", "arctan"], ["linear", "cubic"]],
}
MODULE_NAMES = ["iage", "dye"]


class Interrupted(Exception):
    pass


class FakeState:
    """see the module docstring; `configure(problem)` selects the blocks"""

    __array_priority__ = 100
    model_config_obj = SimpleNamespace(region_cnt=1)
    problem, kinds = "mild", PROBLEMS["mild"]
    calls = 0
    fail_at = None
    members = 1
    A = _A0  # (block [0][0]; kept for the tests that check the solution)
    b = _B0

    @classmethod
    def configure(cls, problem):
        cls.problem, cls.kinds = problem, PROBLEMS[problem]
        cls.model_config_obj = SimpleNamespace(region_cnt=len(cls.kinds[0]))
        cls.calls, cls.fail_at = 0, None

    @classmethod
    def shape(cls):
        return (len(cls.kinds), len(cls.kinds[0]), NVEC)

    @classmethod
    def block_matrix(cls, m, r):
        """block (0, 0) is the matrix of the one-block problems; the others are shifted and rescaled copies"""
        shift = 0.4 * (2 * m + r)
        return _A0 + shift * np.eye(NVEC), _B0 * (1.0 + 0.5 * m - 0.25 * r)

    @classmethod
    def fcn_of(cls, x):
        res = np.empty_like(x)
        for m, row in enumerate(cls.kinds):
            for r, kind in enumerate(row):
                mat, rhs = cls.block_matrix(m, r)
                resid = mat @ x[m, r] - rhs
                if kind == "cubic":
                    res[m, r] = resid + 0.02 * x[m, r] ** 3
                elif kind == "linear":
                    # (scaled like F = x(T) - x(0) of a dissipative model, so that x + F(x) contracts)
                    res[m, r] = -0.1 * resid
                else:
                    # Newton's full step overshoots from x = 1 (the arctangent flattens): the Armijo loop has to damp
                    # it; negative and small in slope, so that the post-Newton fixed-point iteration x + F(x) contracts
                    res[m, r] = -(0.05 * np.arctan(3.0 * resid) + 0.005 * resid)
        return res

    def __init__(self, vals):
        if isinstance(vals, str):
            with open(vals, "rb") as fptr:
                vals = np.load(fptr)
        vals = np.array(vals, dtype=float)
        if vals.ndim == 1:  # one 6-vector: the same start in every block
            vals = np.broadcast_to(vals, self.shape()).copy()
        assert vals.shape == self.shape()
        self.vals = vals
        self.tracer_modules = [SimpleNamespace(name=MODULE_NAMES[m], units="years") for m in range(vals.shape[0])]

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
        res = FakeState(cls.fcn_of(self.vals))
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
        res = np.empty_like(self.vals)
        for m, row in enumerate(self.kinds):
            for r in range(len(row)):
                res[m, r] = self.vals[m, r] / np.diag(self.block_matrix(m, r)[0])
        res = FakeState(res).dump(res_fname, "precond")
        if solver_state is not None:
            solver_state.log_step(step)
        return res

    # reductions: [n_modules, region_cnt]
    def dot_prod(self, other):
        return np.mean(self.vals * other.vals, axis=-1)

    def norm(self):
        return np.sqrt(self.dot_prod(self))

    def mean(self):
        return np.mean(self.vals, axis=-1)

    def mod_gram_schmidt(self, basis_cnt, fname_fcn, quantity):
        n_mod, region_cnt, _ = self.vals.shape
        h = np.zeros((n_mod, basis_cnt, region_cnt))
        for i in range(basis_cnt):
            v = fname_fcn(quantity, i)
            if not isinstance(v, FakeState):  # (the reference hands file names, this package's solver resident states)
                v = FakeState(v)
            h[:, i, :] = self.dot_prod(v)
            self.vals -= h[:, i, :, None] * v.vals
        return h

    # operators with states, floats and [n_modules(, region_cnt)] arrays
    def _s(self, other):
        if isinstance(other, FakeState):
            return other.vals
        arr = np.asarray(other, dtype=float)
        if arr.ndim == 0:
            return float(arr)
        if arr.shape == self.vals.shape[:2]:
            return arr[:, :, None]
        if arr.shape == self.vals.shape[:1]:
            return arr[:, None, None]
        raise ValueError(f"operand of shape {arr.shape}")

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
        return np.ones(self.vals.shape[:2])

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
