"""GPU-resident Newton-Krylov driver.

The loops of the reference's NewtonSolver.step (nk_ooc/newton_solver.py:140-334) and
KrylovSolver.solve (nk_ooc/krylov_solver.py:86-182, left-preconditioned GMRES, Saad alg. 9.4,
x0 = 0) over the device operator surface of this package.  Differences from the reference are in
WHERE the data lives, not in the algorithm:

* the Krylov basis v_j, the preconditioned products w_j and the right-hand side stay in HBM as
  ModelState objects; modified Gram-Schmidt and the linear combinations read them there instead of
  re-reading basis_jj.nc / w_jj.nc for every inner product (j + 1 file reads per iteration in the
  reference);
* Armijo candidates can be evaluated speculatively: `armijo_batch` = k evaluates the factors
  1, 1/2, ..., 2^-(k-1) as k members of ONE batched model-year evaluation and takes the first that
  satisfies the Armijo condition for every (tracer module, region) — identical to the reference's
  sequence whenever all regions accept/reject together (always the case with one region);
  `armijo_batch` = 1 is the reference's sequential per-region halving;
* when `workdir` is given the same files as the reference are written (krylov_NN/precond_fcn_00.nc,
  basis_jj.nc, w_raw_jj.nc, w_jj.nc, krylov_res_jj.nc, increment_NN.nc, iterate_NN.nc, fcn_NN.nc,
  hist_NN.nc), so that baseline_cmp-style comparisons keep working.

The small dense least-squares problem (numpy.linalg.lstsq, krylov_solver.py:168-182) stays on the
host: it is (j+2) x (j+1) per (module, region).
"""

import logging
import os
import tempfile

import numpy as np

from . import distributed, model_state_base, solver_state


def comp_krylov_basis_coeffs(beta, h_mat):
    """least-squares coefficients of the Krylov basis per (tracer module, region)
    (krylov_solver.py:168-182).  beta [n_modules, R], h_mat [n_modules, j+2, j+1, R]"""
    h_shape = h_mat.shape
    coeff = np.zeros((h_shape[0], h_shape[2], h_shape[3]))
    lstsq_rhs = np.zeros(h_shape[1])
    for m in range(h_shape[0]):
        for r in range(h_shape[3]):
            lstsq_rhs[0] = beta[m, r]
            coeff[m, :, r] = np.linalg.lstsq(h_mat[m, :, :, r], lstsq_rhs, rcond=None)[0]
    return coeff


class ProbePreconditioner:
    """Preconditioner built from coloured perturbation probes (the third batched workload of the
    north star): ONE batched function evaluation with n_colours * T * nz members
    (colouring.probe_batch: a unit perturbation of level k, tracer t in every column of one colour)
    gives, per ypos column, the Jacobian block of F with respect to the column itself and to its
    `reach` neighbours (colouring.decode_probes).  M = that block-(tri)diagonal matrix; applying the
    preconditioner is a member-batched banded solve on the device (K4) in column-major ordering.
    For a linear module without lateral processes M is the exact Jacobian of F.  With lateral processes `reach` = 2
    (5 colours) is what makes Newton converge fast: on the refined 125 x 150 grid every Newton step reduces |F| 20 x
    with reach 2 against 2.5 x with reach 1 (measured table), and the reference's own preconditioner
    does not converge there at all.

    Two-dimensional models with ONE tracer module (py_driver_2d)."""

    def __init__(self, iterate, fcn=None, eps=None, reach=1, couple_neighbours=True):
        import torch

        from . import colouring, engine

        if len(iterate.tracer_modules) != 1:
            raise NotImplementedError("ProbePreconditioner handles states with one tracer module")
        tms = iterate.tracer_modules[0]
        if len(tms.cell_shape) != 2:
            raise NotImplementedError("ProbePreconditioner needs a (depth, ypos) grid")
        T, (nz, ny) = tms.tracer_cnt, tms.cell_shape
        x0 = tms.vals[..., 0].cpu().numpy()
        if fcn is None:
            fcn = iterate.comp_fcn(None, None)
        f0 = fcn.tracer_modules[0].vals[..., 0].cpu().numpy()
        if eps is None:
            eps = 1.0e-4 * max(float(np.abs(x0).max()), 1.0e-30)
        colour = colouring.column_colouring(ny, reach)
        probes = colouring.probe_batch(x0, colour, eps)
        B = probes.shape[0]
        batched = type(iterate)("zeros", members=B)
        batched.tracer_modules[0].vals[..., :B] = torch.from_numpy(
            np.ascontiguousarray(np.moveaxis(probes, 0, -1))).cuda()
        # one batched evaluation; inside a process group the probes are sharded over the GPUs and the
        # result columns gathered over NCCL (distributed.sharded_comp_fcn)
        fprobe = np.moveaxis(distributed.sharded_comp_fcn(batched).tracer_modules[0].vals[..., :B].cpu().numpy(), -1, 0)
        jac = colouring.decode_probes(f0, fprobe, colour, eps, reach)  # [ny, 2*reach+1, n, n]
        self.members_probed = B
        n = T * nz
        self._shape = (T, nz, ny)
        # column-major ordering (column slowest): block (j, j+d) sits at rows j*n.., columns (j+d)*n..
        # jac[j, d+reach][:, c] = d F[:, column j+d] / d x[c, column j]  ->  block (row j+d, column j)
        reach_used = reach if couple_neighbours else 0
        kl = ku = (reach_used + 1) * n - 1
        N = ny * n
        ab = np.zeros((kl + ku + 1, N))
        for j in range(ny):
            for d in range(-reach_used, reach_used + 1):
                jj = j + d
                if 0 <= jj < ny:
                    blk = jac[j, d + reach]
                    rows = jj * n + np.arange(n)[:, None]
                    cols = j * n + np.arange(n)[None, :]
                    ab[ku + rows - cols, cols] = blk
        self._factor = engine.BandedFactor(ab, kl, ku)
        self.jac_blocks = jac

    def __call__(self, y):
        """M^-1 y for every member of y"""
        T, nz, ny = self._shape
        res = y._like(clone_vals=False)
        tms = y.tracer_modules[0]
        ldb = tms.vals.shape[-1]
        yp = tms.vals.permute(2, 0, 1, 3).reshape(ny * T * nz, ldb).contiguous()  # layout conversion only
        sol = self._factor.solve(yp, y.members)
        res.tracer_modules[0].vals = sol.reshape(ny, T, nz, ldb).permute(1, 2, 0, 3).contiguous()
        return res


class LaggedPrecond:
    """precond_factory wrapper that keeps a preconditioner for `lag` Newton iterations before it builds the next one
    (lag=None: for ever).  The Jacobian of F does not depend on the iterate for the linear modules (iage, dye_decay:
    F is affine in x), so ONE set of coloured probes serves the whole solve; for the others a lagged preconditioner
    only costs Krylov iterations.  On the refined 125 x 150 grid building the probe preconditioner is half of a
    Newton step (probe batch, band assembly, factorisation: 2.3 of 4.7 s)."""

    def __init__(self, factory, lag=None):
        self._factory, self._lag = factory, lag
        self._cached, self._age = None, 0
        self.built = 0

    def __call__(self, iterate, fcn):
        if self._cached is None or (self._lag is not None and self._age >= self._lag):
            self._cached, self._age = self._factory(iterate, fcn), 0
            self.built += 1
        self._age += 1
        return self._cached


class _MemState:
    """SolverState look-alike for a solve that writes no files (dump=False): nothing is ever 'logged', values live
    in memory.  Lets the solvers below keep ONE control flow — the reference's, with its step log — whether or not
    the work directory is kept (solver_state.py:14-146)."""

    def __init__(self, workdir):
        self._workdir, self._iteration, self._vals = workdir, 0, {}

    def get_workdir(self):
        return self._workdir

    def get_iteration(self):
        return self._iteration

    def inc_iteration(self):
        self._iteration += 1
        return self._iteration

    def log_step(self, stepval, per_iteration=True):  # pylint: disable=unused-argument
        return None

    def step_logged(self, stepval, per_iteration=True):  # pylint: disable=unused-argument
        return False

    def step_was_rewound(self, stepval, per_iteration=True):  # pylint: disable=unused-argument
        return False

    def set_value_saved_state(self, key, value):
        self._vals[key] = value

    def get_value_saved_state(self, key):
        return self._vals[key]


class _SolverBase:
    """work directory, step log and stats file of a solver (solver_base.py:11-191)"""

    def __init__(self, name, workdir, iterate, var_table, dump, resume, rewind):
        self._solver_name = name
        self._workdir = workdir
        self._dump = dump
        os.makedirs(workdir, exist_ok=True)
        if not dump:
            if resume or rewind:
                raise ValueError("resume needs the files of a dumped solve")
            self._solver_state, self._stats = _MemState(workdir), None
            return
        self._solver_state = solver_state.SolverState(name, workdir, resume, rewind)
        cfg = type(iterate).model_config_obj
        mods = [(tms.name, getattr(tms, "units", None)) for tms in iterate.tracer_modules]
        stats_fname = os.path.join(workdir, f"{name}_stats.nc")
        created = self._solver_state.step_logged(f"_create_stats_file {stats_fname}", per_iteration=False)
        self._stats = solver_state.StatsFile(name, workdir, cfg.region_cnt, mods, var_table, resume=created)
        self._solver_state.log_step(f"_create_stats_file {stats_fname}", per_iteration=False)

    def _log_stats_vars_defined(self):
        """solver_base.py:120-125 (the variables themselves are defined when the stats file is created)"""
        self._solver_state.log_step(f"define {self._solver_name} solver stats file vars", per_iteration=False)

    def get_iteration(self):
        return self._solver_state.get_iteration()

    def _fname(self, quantity, iteration=None):
        """solver_base.py:49-53; None when no files are written"""
        if not self._dump:
            return None
        iteration = self.get_iteration() if iteration is None else iteration
        return os.path.join(self._workdir, f"{quantity}_{iteration:02}.nc")

    def _put_stats(self, **kwargs):
        """one 'write <key> vals to stats file' step per key, in the order given (solver_base.py:160-191)"""
        if self._stats is None:
            return
        for key, vals in kwargs.items():
            step = f"write {key} vals to stats file"
            if self._solver_state.step_logged(step):
                continue
            self._stats.put(self.get_iteration(), **{key: vals})
            self._solver_state.log_step(step)

    def _put_stats_invariant(self, **kwargs):
        """solver_base.py:127-158"""
        if self._stats is None:
            return
        for key, vals in kwargs.items():
            step = f"write {key} vals to stats file"
            if self._solver_state.step_logged(step, per_iteration=False):
                continue
            self._stats.put_invariant(**{key: vals})
            self._solver_state.log_step(step, per_iteration=False)


class KrylovSolver(_SolverBase):
    """left-preconditioned GMRES for  J(iterate) x = -fcn  with the basis resident in HBM.
    `precond`: optional callable ModelState -> ModelState applying M^-1 (e.g. ProbePreconditioner)
    instead of the model's apply_precond_jacobian.
    With dump=True the work directory carries the reference's Krylov_state.json (step log, beta, h_mat) and
    Krylov_stats.nc; resume / rewind continue an interrupted solve at the step where it stopped: the basis and
    the preconditioned products of the completed iterations are read back from basis_jj.nc / w_jj.nc."""

    def __init__(self, iterate, solverinfo, hist_fname, workdir, max_iter=50, precond=None, dump=True, resume=False,
                 rewind=False):
        super().__init__("Krylov", workdir, iterate, solver_state.KRYLOV_VARS, dump, resume, rewind)
        self._log_stats_vars_defined()
        self._precond = precond
        self._iterate = iterate
        self._info = solverinfo
        self._max_iter = max_iter
        self.basis, self.w = [], []
        self.beta = None
        self.h_mat = None
        self.precond_resid_norm = []
        # the precond file is needed even when nothing else is kept: the model's preconditioner reads it
        self.precond_fname = os.path.join(workdir, "precond_00.nc")
        if precond is None:
            step = f"ModelStateBase.gen_precond_jacobian {self.precond_fname}"  # model_state_base.py:404-406
            if not self._solver_state.step_logged(step, per_iteration=False):
                iterate.gen_precond_jacobian(hist_fname, self.precond_fname, solver_state=None)
                self._solver_state.log_step(step, per_iteration=False)

    @property
    def iteration(self):
        return self.get_iteration()

    def _state_arg(self):
        """solver_state handed to the model's methods: they log '<method> complete for <res_fname>' steps"""
        return self._solver_state if self._dump else None

    def _apply_precond(self, state, res_fname, caller):
        if self._precond is None:
            return state.apply_precond_jacobian(self.precond_fname, res_fname, self._state_arg())
        step = f"apply_precond_jacobian complete for {res_fname}"
        if self._dump and self._solver_state.step_logged(step):
            return type(self._iterate)(res_fname)
        res = self._precond(state).dump(res_fname, caller)
        self._solver_state.log_step(step)
        return res

    def _rel_tol(self):
        return float(self._info["krylov_rel_tol"])

    def _min_iter(self):
        return int(self._info.get("krylov_min_iter", 0))

    def converged(self, precond_resid_norm):
        """krylov_solver.py:76-84"""
        return (self.get_iteration() >= self._min_iter()) & (precond_resid_norm < self._rel_tol() * self.beta)

    def _resident(self, quantity, ind):
        return self.basis[ind] if quantity == "basis" else self.w[ind]

    def _solve0(self, fcn):
        """step 1 of alg. 9.4: r0 = -M^-1 fcn, beta = ||r0||, v0 = r0 / beta (krylov_solver.py:86-103)"""
        step = "KrylovSolver._solve0"
        cls = type(self._iterate)
        if self._solver_state.step_logged(step, per_iteration=False):
            self.beta = np.asarray(self._solver_state.get_value_saved_state("beta"))
            return cls(self._fname("precond_fcn", 0))
        caller = f"{type(self).__name__}._solve0"
        precond_fcn = self._apply_precond(fcn, self._fname("precond_fcn"), caller)
        self.beta = precond_fcn.norm()
        fcn.log_vals("beta", self.beta)
        self._put_stats_invariant(precond_rhs_norm=self.beta)
        (-precond_fcn / self.beta).dump(self._fname("basis"), caller)
        self._solver_state.set_value_saved_state("beta", np.asarray(self.beta))
        self._solver_state.log_step(step, per_iteration=False)
        return precond_fcn

    def solve(self, res_fname, fcn, dump=None):  # pylint: disable=unused-argument
        """krylov_solver.py:105-165.  (`dump` is accepted for compatibility; the constructor decides.)"""
        logger = logging.getLogger(__name__)
        caller = f"{type(self).__name__}.solve"
        cls = type(self._iterate)
        resumed = self._dump and self._solver_state.step_logged("KrylovSolver._solve0", per_iteration=False)
        precond_fcn = self._solve0(fcn)
        if resumed:
            # basis vectors and preconditioned products of the completed iterations, back into HBM
            j_done = self.get_iteration()
            self.basis = [cls(self._fname("basis", j)) for j in range(j_done)]
            self.w = [cls(self._fname("w", j)) for j in range(j_done)]
            if j_done > 0:
                self.h_mat = np.asarray(self._solver_state.get_value_saved_state("h_mat"))
            self.basis.append(cls(self._fname("basis")))
        else:
            self.basis.append(-precond_fcn / self.beta)
        n_mod, region_cnt = self.beta.shape[0], self.beta.shape[1]
        while True:
            j_val = self.get_iteration()
            h_mat = np.zeros((n_mod, j_val + 2, j_val + 1, region_cnt))
            if j_val > 0:
                h_mat[:, :-1, :-1, :] = self.h_mat
            w_raw = self._iterate.comp_jacobian_fcn_state_prod(fcn, self.basis[j_val], self._fname("w_raw"),
                                                               self._state_arg())
            w_j = self._apply_precond(w_raw, self._fname("w"), caller)
            self.w.append(w_j._like())  # un-orthogonalised M^-1 J v_j: needed for the residual below
            h_mat[:, :-1, -1, :] = w_j.mod_gram_schmidt(j_val + 1, self._resident, "basis")
            h_mat[:, -1, -1, :] = w_j.norm()
            w_j /= h_mat[:, -1, -1, :]
            self.h_mat = h_mat
            self._solver_state.set_value_saved_state("h_mat", h_mat)
            coeff = comp_krylov_basis_coeffs(self.beta, h_mat)
            self._iterate.log_vals("KrylovCoeff", coeff)
            res = model_state_base.lin_comb(cls, coeff, self._resident, "basis")
            res.dump(self._fname("krylov_res", j_val), caller)
            precond_resid = model_state_base.lin_comb(cls, coeff, self._resident, "w")
            precond_resid += precond_fcn
            resid_norm = precond_resid.norm()
            self._iterate.log_vals("precond_resid", resid_norm)
            self.precond_resid_norm.append(resid_norm)
            self._put_stats(precond_resid_norm=resid_norm)
            logger.info("Krylov iteration %d: precond_resid_norm/beta = %s", j_val, resid_norm / self.beta)
            self._solver_state.inc_iteration()
            if self.converged(resid_norm).all():
                logger.info("Krylov convergence criterion satisfied")
                break
            if self.get_iteration() >= self._max_iter:
                raise RuntimeError("number of maximum Krylov iterations exceeded")
            self.basis.append(w_j.dump(self._fname("basis"), caller))
        return res.dump(res_fname, caller)


class NewtonSolver(_SolverBase):
    """Newton's method with Armijo damping and post-Newton fixed-point iterations (newton_solver.py:22-334) on a
    device-resident iterate.

    With dump=True (the default) the work directory receives the reference's files and its Newton_state.json —
    the SAME step strings in the same order as nk_ooc/newton_solver.py + solver_base.py + stats_file.py log them
    (baselines/ci_*/Newton_state.json are reproduced verbatim) — so that `resume=True` continues an interrupted
    solve at the step where it stopped and `rewind=True` redoes the last logged step (solver_state.py:36-45,
    91-98): every intermediate the reference reads back (increment, Armijo candidates, fixed-point iterates) is
    read back here too instead of being recomputed.  With dump=False nothing is written and nothing can be resumed;
    the control flow is the same."""

    def __init__(self, iterate, solverinfo, workdir=None, armijo_batch=1, dump=True, precond_factory=None,
                 resume=False, rewind=False):
        """precond_factory: optional callable (iterate, fcn) -> preconditioner callable, called once per
        Newton iteration (e.g. ProbePreconditioner)."""
        self._precond_factory = precond_factory
        self._info = dict(solverinfo)
        workdir = workdir or tempfile.mkdtemp(prefix="nkb200_newton_")
        super().__init__("Newton", workdir, iterate, solver_state.NEWTON_VARS, dump, resume, rewind)
        self._armijo_batch = int(armijo_batch)
        state = self._solver_state
        caller = f"{type(self).__name__}.__init__"
        step0 = "Newton iterate 0 written"
        if state.step_logged(step0, per_iteration=False):
            iterate = type(iterate)(self._fname("iterate"))
        else:
            iterate.copy_real_tracers_to_shadow_tracers().dump(self._fname("iterate"), caller)
            state.log_step(step0, per_iteration=False)
        self._log_stats_vars_defined()
        self._iterate = iterate
        self._fcn = iterate.comp_fcn(self._fname("fcn"), self._state_arg(), self._hist_fname("hist"))
        self._put_stats(iterate=self._iterate, fcn=self._fcn)
        self._put_hist_stats(self._iterate)
        self.history = []  # per iteration: dict(fcn_norm, iterate_norm, krylov_iterations, armijo_factor, ...)
        self._record()

    # ---- bookkeeping --------------------------------------------------------------------
    @property
    def iteration(self):
        return self.get_iteration()

    def _state_arg(self):
        return self._solver_state if self._dump else None

    def _hist_fname(self, quantity, iteration=None):
        """hist files are written even by a solve that keeps nothing else: the next preconditioner reads them"""
        iteration = self.get_iteration() if iteration is None else iteration
        return os.path.join(self._workdir, f"{quantity}_{iteration:02}.nc")

    def _put_hist_stats(self, state):
        """the model's own statistics of this iteration's hist file (newton_solver.py:52-58,330 call the
        three methods of the operator surface)"""
        if self._stats is None:
            return
        hist_fname = self._hist_fname("hist")
        state.def_stats_vars(self._stats, hist_fname, self._solver_state)
        state.put_stats_vars_iteration_invariant(self._stats, hist_fname, self._solver_state)
        state.put_stats_vars(self._stats, hist_fname, self._solver_state)

    def _record(self, **extra):
        rec = {"iteration": self.iteration, "fcn_norm": self._fcn.norm(), "iterate_norm": self._iterate.norm()}
        rec.update(extra)
        self.history.append(rec)
        logging.getLogger(__name__).info("Newton iteration %02d: |fcn|/|iterate| = %s", self.iteration,
                                         rec["fcn_norm"] / np.where(rec["iterate_norm"] == 0, 1, rec["iterate_norm"]))

    @property
    def iterate(self):
        return self._iterate

    @property
    def fcn(self):
        return self._fcn

    def converged(self):
        """newton_solver.py:133-138"""
        rel_tol = float(self._info["newton_rel_tol"])
        min_iter = int(self._info.get("newton_min_iter", 0))
        return (self.iteration >= min_iter) & (self._fcn.norm() < rel_tol * self._iterate.norm())

    def converged_flat(self):
        return self.converged().all()

    # ---- one Newton step ------------------------------------------------------------------
    def _comp_increment(self):
        """(d fcn / d iterate) increment = -fcn (newton_solver.py:140-181); returns (increment, krylov solver or
        None when the increment of an interrupted solve was read back)"""
        state = self._solver_state
        done_step = "_comp_increment complete"
        if state.step_logged(done_step):
            return type(self._iterate)(self._fname("increment")), None
        krylov_dir = os.path.join(self._workdir, f"krylov_{self.iteration:02}")
        step = "KrylovSolver instantiated"
        rewind = state.step_was_rewound(step)
        resume = rewind or state.step_logged(step)
        precond = None if self._precond_factory is None else self._precond_factory(self._iterate, self._fcn)
        krylov = KrylovSolver(self._iterate, self._info, self._hist_fname("hist"), krylov_dir, precond=precond,
                              dump=self._dump, resume=resume, rewind=rewind)
        state.log_step(step)
        increment = krylov.solve(self._fname("increment"), self._fcn)
        self._put_stats(Krylov_iterations=krylov.iteration, increment=increment)
        state.log_step(done_step)
        return increment, krylov

    def _armijo_init(self):
        """newton_solver.py:183-189"""
        state = self._solver_state
        step = "NewtonSolver._armijo_init"
        if not state.step_logged(step):
            state.set_value_saved_state("armijo_ind", 0)
            state.set_value_saved_state("armijo_factor", np.where(self.converged(), 0.0, 1.0))
            state.log_step(step)

    def _comp_next_iterate(self, increment):
        """Armijo damping, Eq. (A.1) of Kelley 2003 (newton_solver.py:191-258); returns
        (prov, prov_fcn, armijo_factor, armijo_ind)"""
        state = self._solver_state
        self._armijo_init()
        armijo_ind = int(state.get_value_saved_state("armijo_ind"))
        armijo_factor = np.asarray(state.get_value_saved_state("armijo_factor"), dtype=np.float64)
        done_step = "_comp_next_iterate complete"
        cls = type(self._iterate)
        if state.step_logged(done_step):
            return (cls(self._fname(f"prov_Armijo_{armijo_ind:02}")), cls(self._fname(f"prov_fcn_Armijo_{armijo_ind:02}")),
                    armijo_factor, armijo_ind)
        caller = f"{type(self).__name__}._comp_next_iterate"
        alpha = 1.0e-4
        fcn_norm = self._fcn.norm()
        if self._armijo_batch > 1 and armijo_ind == 0:
            # speculative: k candidates as k members of ONE batched evaluation
            k = self._armijo_batch
            factors = [armijo_factor * 0.5 ** i for i in range(k)]
            provs = [self._iterate + f * increment for f in factors]
            batched_fcn = distributed.sharded_comp_fcn(cls.from_members(provs))
            norms = batched_fcn.norm()  # [n_modules, R, k]
            for i in range(k):
                cond = (factors[i] == 0.0) | (norms[..., i] <= (1.0 - alpha * factors[i]) * fcn_norm)
                if cond.all():
                    state.set_value_saved_state("armijo_ind", i)
                    state.set_value_saved_state("armijo_factor", factors[i])
                    prov = provs[i].dump(self._fname(f"prov_Armijo_{i:02}"), caller)
                    prov_fcn = batched_fcn.member(i).dump(self._fname(f"prov_fcn_Armijo_{i:02}"), caller)
                    state.log_step(done_step)
                    self._put_stats(Armijo_factor=factors[i])
                    self._armijo_hist = None  # a batched evaluation leaves no hist file of the accepted candidate
                    return prov, prov_fcn, factors[i], i
            armijo_factor, armijo_ind = factors[-1] * 0.5, k
            state.set_value_saved_state("armijo_ind", armijo_ind)
            state.set_value_saved_state("armijo_factor", armijo_factor)
        while True:
            prov = self._iterate + armijo_factor * increment
            prov.dump(self._fname(f"prov_Armijo_{armijo_ind:02}"), caller)
            hist = self._hist_fname(f"prov_hist_Armijo_{armijo_ind:02}")
            prov_fcn = prov.comp_fcn(self._fname(f"prov_fcn_Armijo_{armijo_ind:02}"), self._state_arg(), hist)
            # only the latest Armijo hist file is kept
            prev = self._hist_fname(f"prov_hist_Armijo_{(armijo_ind - 1):02}")
            if armijo_ind > 0 and os.path.exists(prev):
                os.remove(prev)
            self._armijo_hist = hist
            prov_fcn_norm = prov_fcn.norm()
            increment.log_vals(["ArmijoFactor", "fcn_norm", "prov_fcn_norm"],
                               np.stack((armijo_factor, fcn_norm, prov_fcn_norm)))
            armijo_cond = (armijo_factor == 0.0) | (prov_fcn_norm <= (1.0 - alpha * armijo_factor) * fcn_norm)
            if armijo_cond.all():
                state.log_step(done_step)
                self._put_stats(Armijo_factor=armijo_factor)
                return prov, prov_fcn, armijo_factor, armijo_ind
            armijo_factor = np.where(armijo_cond, armijo_factor, 0.5 * armijo_factor)
            armijo_ind += 1
            state.set_value_saved_state("armijo_ind", armijo_ind)
            state.set_value_saved_state("armijo_factor", armijo_factor)
            if armijo_ind > 10:
                raise RuntimeError("Armijo_ind exceeds limit")

    def step(self):
        """newton_solver.py:260-334"""
        if self.iteration >= int(self._info["newton_max_iter"]):
            raise RuntimeError("number of maximum Newton iterations exceeded")
        caller = f"{type(self).__name__}.step"
        state = self._solver_state
        cls = type(self._iterate)
        n_fp = int(self._info.get("post_newton_fp_iter", 0))
        extra = {}
        increment = None
        step = "fp iterations started"
        if not state.step_logged(step):
            increment, krylov = self._comp_increment()
            scalef = increment.apply_limiter(self._iterate)
            self._put_stats(increment_scalef=scalef)
            self._armijo_hist = None
            prov, prov_fcn, armijo_factor, armijo_ind = self._comp_next_iterate(increment)
            extra = dict(increment_scalef=scalef, armijo_factor=armijo_factor, armijo_ind=armijo_ind)
            if krylov is not None:
                extra.update(krylov_iterations=krylov.iteration, krylov_precond_resid_norm=krylov.precond_resid_norm,
                             krylov_beta=krylov.beta)
            fp_iter = 0
            state.set_value_saved_state("fp_iter", fp_iter)
            prov.copy_shadow_tracers_to_real_tracers()
            prov.dump(self._fname(f"prov_fp_{fp_iter:02}"), caller)
            # the function value after the shadow tracers were copied; without shadow tracers it is the accepted
            # Armijo candidate's, whose hist file becomes the first fixed-point hist file (newton_solver.py:283-303)
            armijo_hist = self._armijo_hist
            if armijo_hist is None and self._dump:
                armijo_hist = self._hist_fname(f"prov_hist_Armijo_{armijo_ind:02}")
            have_hist = armijo_hist is not None and os.path.exists(armijo_hist)
            fp_hist = self._hist_fname(f"prov_hist_fp_{fp_iter:02}")
            if prov.shadow_tracers_on() or (n_fp == 0 and not have_hist):
                prov_fcn = prov.comp_fcn(self._fname(f"prov_fcn_fp_{fp_iter:02}"), self._state_arg(), fp_hist)
                if have_hist:
                    os.remove(armijo_hist)
            else:
                prov_fcn.dump(self._fname(f"prov_fcn_fp_{fp_iter:02}"), caller)
                if have_hist:
                    os.replace(armijo_hist, fp_hist)
            state.log_step(step)
        else:
            fp_iter = int(state.get_value_saved_state("fp_iter"))
            prov = cls(self._fname(f"prov_fp_{fp_iter:02}"))
            prov_fcn = cls(self._fname(f"prov_fcn_fp_{fp_iter:02}"))
        if n_fp == 0 and fp_iter == 0:
            # (the reference only advances the iteration inside its fixed-point loop; without fixed-point
            # iterations the accepted Armijo candidate IS the new iterate and its function value and hist file
            # — the input of the next preconditioner — exist already: no further model year)
            fp_hist = self._hist_fname("prov_hist_fp_00")
            state.inc_iteration()
            prov.dump(self._fname("iterate"), caller)
            prov_fcn.dump(self._fname("fcn"), caller)
            if os.path.exists(fp_hist):
                os.replace(fp_hist, self._hist_fname("hist"))
            state.log_step(f"comp_fcn complete for {self._fname('fcn')}")
            fp_iter = 1
            state.set_value_saved_state("fp_iter", fp_iter)
        while fp_iter < n_fp:
            step = f"prov updated for fp iteration {fp_iter:02}"
            if not state.step_logged(step):
                prov += prov_fcn
                prov.copy_shadow_tracers_to_real_tracers()
                prov.dump(self._fname(f"prov_fp_{(fp_iter + 1):02}"), caller)
                state.log_step(step)
            else:
                prov = cls(self._fname(f"prov_fp_{(fp_iter + 1):02}"))
            if fp_iter + 1 < n_fp:
                res_fname = self._fname(f"prov_fcn_fp_{(fp_iter + 1):02}")
                hist_fname = self._hist_fname(f"prov_hist_fp_{(fp_iter + 1):02}")
            else:
                state.inc_iteration()
                prov.dump(self._fname("iterate"), caller)
                res_fname = self._fname("fcn")
                hist_fname = self._hist_fname("hist")
            prov_fcn = prov.comp_fcn(res_fname, self._state_arg(), hist_fname)
            fp_iter += 1
            state.set_value_saved_state("fp_iter", fp_iter)
        self._iterate, self._fcn = prov, prov_fcn
        self._put_stats(iterate=self._iterate, fcn=self._fcn)
        if self._stats is not None:
            self._iterate.put_stats_vars(self._stats, self._hist_fname("hist"), self._solver_state)
        self._record(**extra)
        return increment

    def solve(self):
        """iterate until converged (nk_driver.py main loop)"""
        while not self.converged_flat():
            self.step()
        return self._iterate
