"""Newton-Krylov on a large py_driver_2d grid from gen_init_iterate: the first Newton step with the reference's
preconditioner (with the solver's log, Armijo candidates and the true linear residual), then up to four Newton steps
with the probe preconditioner:   python scripts/nk_large_grid.py 125 150 11.8 [krylov_rel_tol]"""
import logging, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "newton-krylov_ooc_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from nk_ooc_b200.py_driver_2d.model_state import ModelState
from nk_ooc_b200.py_driver_2d.setup_solver import gen_grid_vars_file
from nk_ooc_b200.solver import KrylovSolver, NewtonSolver
from test_gpu_model_state import _modelinfo

logging.basicConfig(level=logging.INFO, format="%(message)s")
nz, ny, ratio = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
tmp = tempfile.mkdtemp()
info = _modelinfo(tmp, nz, ny)
info["depth_delta_ratio_max"] = ratio
gen_grid_vars_file(info)
ModelState.configure(info)
it = ModelState("gen_init_iterate")
pd_info = {"newton_rel_tol": "1.0e-5", "newton_max_iter": "8", "post_newton_fp_iter": "1",
           "krylov_rel_tol": sys.argv[4] if len(sys.argv) > 4 else "0.01"}
PROBE_ONLY = os.environ.get("NK_PROBE_ONLY") == "1"
if PROBE_ONLY:
    import cProfile, pstats, time
    from nk_ooc_b200.solver import ProbePreconditioner
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pr = cProfile.Profile(); pr.enable()
    solver2 = NewtonSolver(ModelState("gen_init_iterate"), pd_info, workdir=os.path.join(tmp, "work2"), dump=False,
                           precond_factory=lambda itr, fcn: ProbePreconditioner(itr, fcn))
    solver2.step()
    torch.cuda.synchronize(); pr.disable()
    print(f"one probe-preconditioned Newton step: {time.perf_counter()-t0:.2f} s, Krylov iterations "
          f"{solver2.history[-1].get('krylov_iterations')}")
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
    sys.exit(0)
solver = NewtonSolver(it, pd_info, workdir=os.path.join(tmp, "work"), dump=True)
print("|F| / |x|", solver.fcn.norm() / solver.iterate.norm())
inc, kr = solver._comp_increment()
print("krylov iterations", kr.iteration, "beta", kr.beta, "resid", kr.precond_resid_norm)
for a in (1.0, 0.5, 0.1):
    prov = solver.iterate + a * inc
    print("armijo factor", a, "|F(prov)|", prov.comp_fcn(None, None).norm(), "|F|", solver.fcn.norm(), flush=True)
# is the increment a solution of J inc = -F ?  (F is affine in x for iage: exact finite differences)
jv = solver.iterate.comp_jacobian_fcn_state_prod(solver.fcn, inc, None, None)
print("|J inc + F| / |F|", (jv + solver.fcn).norm() / solver.fcn.norm())

# the probe preconditioner (block-tridiagonal Jacobian of F from ONE batched evaluation of coloured probes)
import time
from nk_ooc_b200.solver import ProbePreconditioner
torch.cuda.synchronize(); t0 = time.perf_counter()
from nk_ooc_b200.solver import LaggedPrecond
lag = os.environ.get("NK_PRECOND_LAG")  # "inf": one set of probes for the whole solve (exact for the affine iage module)
reach = int(os.environ.get("NK_PROBE_REACH", "1"))  # columns of coupling kept on each side (2*reach + 1 colours)
fac = lambda itr, fcn: ProbePreconditioner(itr, fcn, reach=reach)
if lag:
    fac = LaggedPrecond(fac, None if lag == "inf" else int(lag))
solver2 = NewtonSolver(ModelState("gen_init_iterate"), pd_info, workdir=os.path.join(tmp, "work2"), dump=False,
                       precond_factory=fac)
n = 0
while not solver2.converged_flat() and n < 4:
    solver2.step(); n += 1
torch.cuda.synchronize()
print(f"probe preconditioner: {n} Newton steps in {time.perf_counter()-t0:.2f} s, Krylov iterations "
      f"{[r.get('krylov_iterations') for r in solver2.history[1:]]}, |F|/|x| "
      f"{[float((r['fcn_norm']/r['iterate_norm']).max()) for r in solver2.history]}")
