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
    For a linear module without lateral processes M is the exact Jacobian of F.

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


class KrylovSolver:
    """left-preconditioned GMRES for  J(iterate) x = -fcn  with the basis resident in HBM.
    `precond`: optional callable ModelState -> ModelState applying M^-1 (e.g. ProbePreconditioner)
    instead of the model's apply_precond_jacobian."""

    def __init__(self, iterate, solverinfo, hist_fname, workdir, max_iter=50, precond=None):
        self._precond = precond
        self._iterate = iterate
        self._info = solverinfo
        self._workdir = workdir
        self._max_iter = max_iter
        self.iteration = 0
        self.basis, self.w = [], []
        self.beta = None
        self.h_mat = None
        self.precond_resid_norm = []
        os.makedirs(workdir, exist_ok=True)
        self._state = self._stats = None
        self.precond_fname = self._fname("precond", 0)
        if precond is None:
            iterate.gen_precond_jacobian(hist_fname, self.precond_fname, solver_state=None)

    def _apply_precond(self, state, res_fname, caller):
        if self._precond is None:
            return state.apply_precond_jacobian(self.precond_fname, res_fname, None)
        return self._precond(state).dump(res_fname, caller)

    def _fname(self, quantity, iteration=None):
        iteration = self.iteration if iteration is None else iteration
        return os.path.join(self._workdir, f"{quantity}_{iteration:02}.nc")

    def _rel_tol(self):
        return float(self._info["krylov_rel_tol"])

    def _min_iter(self):
        return int(self._info.get("krylov_min_iter", 0))

    def converged(self, precond_resid_norm):
        """krylov_solver.py:76-84"""
        return (self.iteration >= self._min_iter()) & (precond_resid_norm < self._rel_tol() * self.beta)

    def _resident(self, quantity, ind):
        return self.basis[ind] if quantity == "basis" else self.w[ind]

    def solve(self, res_fname, fcn, dump=True):
        logger = logging.getLogger(__name__)
        caller = f"{type(self).__name__}.solve"
        fn = self._fname if dump else (lambda *a, **k: None)
        # step 1 of alg. 9.4: r0 = -M^-1 fcn, beta = ||r0||, v0 = r0 / beta
        if dump:
            # Krylov_state.json (beta, h_mat as the reference saves them, krylov_solver.py:101,136) and
            # Krylov_stats.nc in the Krylov work directory
            cfg = type(self._iterate).model_config_obj
            mods = [(tms.name, getattr(tms, "units", None)) for tms in self._iterate.tracer_modules]
            self._state = solver_state.SolverState("Krylov", self._workdir)
            self._stats = solver_state.StatsFile("Krylov", self._workdir, cfg.region_cnt, mods, solver_state.KRYLOV_VARS)
        precond_fcn = self._apply_precond(fcn, fn("precond_fcn"), caller)
        self.beta = precond_fcn.norm()
        if self._stats is not None:
            self._stats.put_invariant(precond_rhs_norm=self.beta)
            self._state.set_value_saved_state("beta", np.asarray(self.beta))
        self.basis.append((-precond_fcn / self.beta).dump(fn("basis"), caller))
        n_mod, region_cnt = self.beta.shape[0], self.beta.shape[1]
        while True:
            j_val = self.iteration
            h_mat = np.zeros((n_mod, j_val + 2, j_val + 1, region_cnt))
            if j_val > 0:
                h_mat[:, :-1, :-1, :] = self.h_mat
            w_raw = self._iterate.comp_jacobian_fcn_state_prod(fcn, self.basis[j_val], fn("w_raw"), None)
            w_j = self._apply_precond(w_raw, fn("w"), caller)
            self.w.append(w_j._like())  # un-orthogonalised M^-1 J v_j: needed for the residual below
            h_mat[:, :-1, -1, :] = w_j.mod_gram_schmidt(j_val + 1, self._resident, "basis")
            h_mat[:, -1, -1, :] = w_j.norm()
            w_j /= h_mat[:, -1, -1, :]
            self.h_mat = h_mat
            coeff = comp_krylov_basis_coeffs(self.beta, h_mat)
            res = model_state_base.lin_comb(type(self._iterate), coeff, self._resident, "basis")
            res.dump(fn("krylov_res", j_val), caller)
            precond_resid = model_state_base.lin_comb(type(self._iterate), coeff, self._resident, "w")
            precond_resid += precond_fcn
            resid_norm = precond_resid.norm()
            self.precond_resid_norm.append(resid_norm)
            if self._stats is not None:
                self._state.set_value_saved_state("h_mat", h_mat)
                self._stats.put(j_val, precond_resid_norm=resid_norm)
                self._state.inc_iteration()
            logger.info("Krylov iteration %d: precond_resid_norm/beta = %s", j_val, resid_norm / self.beta)
            self.iteration += 1
            if self.converged(resid_norm).all():
                logger.info("Krylov convergence criterion satisfied")
                break
            if self.iteration >= self._max_iter:
                raise RuntimeError("number of maximum Krylov iterations exceeded")
            self.basis.append(w_j.dump(fn("basis"), caller))
        return res.dump(res_fname, caller)


class NewtonSolver:
    """Newton's method with Armijo damping and post-Newton fixed-point iterations
    (newton_solver.py:22-334) on a device-resident iterate"""

    def __init__(self, iterate, solverinfo, workdir=None, armijo_batch=1, dump=True, precond_factory=None,
                 resume=False):
        """precond_factory: optional callable (iterate, fcn) -> preconditioner callable, called once per
        Newton iteration (e.g. ProbePreconditioner).
        With dump=True the work directory also receives the reference's Newton_state.json (iteration
        counter + step log, solver_state.py) and Newton_stats.nc (stats_file.py); resume=True continues
        from them: the iterate (and, if its evaluation had completed, fcn) of the logged iteration are
        read back instead of recomputed (newton_solver.py:30-47,140-172)."""
        self._precond_factory = precond_factory
        self._info = dict(solverinfo)
        self._workdir = workdir or tempfile.mkdtemp(prefix="nkb200_newton_")
        os.makedirs(self._workdir, exist_ok=True)
        self._armijo_batch = int(armijo_batch)
        self._dump = dump
        if resume and not dump:
            raise ValueError("resume needs the files of a dumped solve")
        self.iteration = 0
        self._state = solver_state.SolverState("Newton", self._workdir, resume=resume) if dump else None
        self._stats = None
        if dump:
            cfg = type(iterate).model_config_obj
            mods = [(tms.name, getattr(tms, "units", None)) for tms in iterate.tracer_modules]
            self._stats = solver_state.StatsFile("Newton", self._workdir, cfg.region_cnt, mods,
                                                 solver_state.NEWTON_VARS, resume=resume)
            self.iteration = self._state.get_iteration()
        caller = f"{type(self).__name__}.__init__"
        step0 = "Newton iterate 0 written"
        if resume and self._state.step_logged(step0, per_iteration=False):
            iterate = type(iterate)(self._fname("iterate"))
        else:
            iterate.dump(self._fname("iterate") if dump else None, caller)
            if dump:
                self._state.log_step(step0, per_iteration=False)
        self._iterate = iterate
        fcn_step = f"comp_fcn complete for {self._fname('fcn')}"
        if resume and self._state.step_logged(fcn_step):
            self._fcn = type(iterate)(self._fname("fcn"))
        else:
            self._fcn = iterate.comp_fcn(self._fname("fcn") if dump else None, None, self._fname("hist"))
            if dump:
                self._state.log_step(fcn_step)
                self._stats.put(self.iteration, iterate=self._iterate, fcn=self._fcn)
                self._put_hist_stats(self._iterate)
        self.history = []  # per iteration: dict(fcn_norm, iterate_norm, krylov_iterations, armijo_factor, ...)
        self._record()

    # ---- bookkeeping --------------------------------------------------------------------
    def _fname(self, quantity, iteration=None):
        iteration = self.iteration if iteration is None else iteration
        return os.path.join(self._workdir, f"{quantity}_{iteration:02}.nc")

    def _put_hist_stats(self, state):
        """the model's own statistics of this iteration's hist file (newton_solver.py:52-58,330 call the
        three methods of the operator surface)"""
        hist_fname = self._fname("hist")
        if not os.path.exists(hist_fname):
            return
        state.def_stats_vars(self._stats, hist_fname, None)
        state.put_stats_vars_iteration_invariant(self._stats, hist_fname, None)
        names, weights = state._stats_names_and_weights()
        self._stats.put_hist_stats(self.iteration, hist_fname, names, weights)

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
        krylov_dir = os.path.join(self._workdir, f"krylov_{self.iteration:02}")
        precond = None if self._precond_factory is None else self._precond_factory(self._iterate, self._fcn)
        krylov = KrylovSolver(self._iterate, self._info, self._fname("hist"), krylov_dir, precond=precond)
        increment = krylov.solve(self._fname("increment") if self._dump else None, self._fcn, dump=self._dump)
        return increment, krylov

    def _armijo_candidates(self, increment, armijo_factor):
        """prov = iterate + factor*increment and F(prov) for one factor array [n_modules, R]"""
        prov = self._iterate + armijo_factor * increment
        return prov, prov.comp_fcn(None, None, self._fname("prov_hist_Armijo"))

    def _comp_next_iterate(self, increment):
        """Armijo damping, Eq. (A.1) of Kelley 2003 (newton_solver.py:183-258)"""
        alpha = 1.0e-4
        armijo_factor = np.where(self.converged(), 0.0, 1.0)
        fcn_norm = self._fcn.norm()
        armijo_ind = 0
        if self._armijo_batch > 1:
            # speculative: k candidates as k members of one batched evaluation
            k = self._armijo_batch
            factors = [armijo_factor * 0.5 ** i for i in range(k)]
            provs = [self._iterate + f * increment for f in factors]
            batched_fcn = distributed.sharded_comp_fcn(type(self._iterate).from_members(provs))
            norms = batched_fcn.norm()  # [n_modules, R, k]
            for i in range(k):
                cond = (factors[i] == 0.0) | (norms[..., i] <= (1.0 - alpha * factors[i]) * fcn_norm)
                if cond.all():
                    return provs[i], batched_fcn.member(i), factors[i], i
            armijo_factor, armijo_ind = factors[-1] * 0.5, k
        while True:
            prov, prov_fcn = self._armijo_candidates(increment, armijo_factor)
            prov_fcn_norm = prov_fcn.norm()
            armijo_cond = (armijo_factor == 0.0) | (prov_fcn_norm <= (1.0 - alpha * armijo_factor) * fcn_norm)
            if armijo_cond.all():
                return prov, prov_fcn, armijo_factor, armijo_ind
            armijo_factor = np.where(armijo_cond, armijo_factor, 0.5 * armijo_factor)
            armijo_ind += 1
            if armijo_ind > 10:
                raise RuntimeError("Armijo_ind exceeds limit")

    def step(self):
        """newton_solver.py:260-334"""
        if self.iteration >= int(self._info["newton_max_iter"]):
            raise RuntimeError("number of maximum Newton iterations exceeded")
        caller = f"{type(self).__name__}.step"
        increment, krylov = self._comp_increment()
        scalef = increment.apply_limiter(self._iterate)
        prov, prov_fcn, armijo_factor, armijo_ind = self._comp_next_iterate(increment)
        prov.copy_shadow_tracers_to_real_tracers()
        if prov.shadow_tracers_on():
            prov_fcn = prov.comp_fcn(None, None, self._fname("prov_hist_fp"))
        n_fp = int(self._info.get("post_newton_fp_iter", 0))
        if n_fp == 0:
            # the accepted Armijo candidate IS the new iterate: its function value and its hist file (the input of
            # the next iteration's preconditioner) exist already — no further model year (newton_solver.py:283-291
            # renames the Armijo hist file the same way).  Only the speculative batched Armijo step and shadow
            # tracers leave no hist file of the accepted candidate behind.
            armijo_hist = self._fname("prov_hist_Armijo")
            have_hist = self._armijo_batch <= 1 and not prov.shadow_tracers_on() and os.path.exists(armijo_hist)
            self.iteration += 1
            prov.dump(self._fname("iterate") if self._dump else None, caller)
            if have_hist:
                os.replace(armijo_hist, self._fname("hist"))
                prov_fcn.dump(self._fname("fcn") if self._dump else None, caller)
            else:
                prov_fcn = prov.comp_fcn(self._fname("fcn") if self._dump else None, None, self._fname("hist"))
        for fp_iter in range(n_fp):
            prov += prov_fcn
            prov.copy_shadow_tracers_to_real_tracers()
            if fp_iter + 1 < n_fp:
                prov_fcn = prov.comp_fcn(None, None, self._fname("prov_hist_fp"))
            else:
                self.iteration += 1
                prov.dump(self._fname("iterate") if self._dump else None, caller)
                prov_fcn = prov.comp_fcn(self._fname("fcn") if self._dump else None, None, self._fname("hist"))
        if self._dump:
            # statistics of the iteration that ends here, then the new iteration's iterate / fcn
            # (newton_solver.py:262-329; the iteration counter was advanced above)
            prev = self.iteration - 1
            self._stats.put(prev, increment=increment, Krylov_iterations=krylov.iteration,
                            increment_scalef=scalef, Armijo_factor=armijo_factor)
            self._state.inc_iteration()
            self._state.log_step(f"comp_fcn complete for {self._fname('fcn')}")
            self._stats.put(self.iteration, iterate=prov, fcn=prov_fcn)
            self._put_hist_stats(prov)
        self._iterate, self._fcn = prov, prov_fcn
        self._record(krylov_iterations=krylov.iteration, krylov_precond_resid_norm=krylov.precond_resid_norm,
                     krylov_beta=krylov.beta, increment_scalef=scalef, armijo_factor=armijo_factor,
                     armijo_ind=armijo_ind)
        return increment

    def solve(self):
        """iterate until converged (nk_driver.py main loop)"""
        while not self.converged_flat():
            self.step()
        return self._iterate
