"""wall-time profile of Newton-Krylov iterations of py_driver_2d iage (single states): the 30 x 30 CI grid from the
CI initial iterate, or any grid from gen_init_iterate:
python scripts/nk_profile.py [dump] [nz ny ratio]      e.g.  python scripts/nk_profile.py nodump 125 150 11.8"""
import cProfile, os, pstats, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "newton-krylov_ooc_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from nk_ooc_b200.py_driver_2d.model_state import ModelState
from nk_ooc_b200.py_driver_2d.setup_solver import gen_grid_vars_file
from nk_ooc_b200.solver import NewtonSolver
from test_gpu_model_state import _modelinfo

dump = len(sys.argv) > 1 and sys.argv[1] == "dump"
tmp = tempfile.mkdtemp()
big = len(sys.argv) > 4
nz, ny = (int(sys.argv[2]), int(sys.argv[3])) if big else (30, 30)
info = _modelinfo(tmp, nz, ny)
if big:
    info["depth_delta_ratio_max"] = sys.argv[4]
gen_grid_vars_file(info)
ModelState.configure(info)
if big:
    it = ModelState("gen_init_iterate")
else:
    base = np.load(os.path.join(ROOT, "tests", "golden", "baselines.npz"))
    pre = "ci_py_driver_2d_iage/init_iterate"
    it = ModelState({"iage": base[pre + "/iage"], "iage_slow_rest": base[pre + "/iage_slow_rest"]})
pd_info = {"newton_rel_tol": "1.0e-5", "newton_max_iter": "8", "post_newton_fp_iter": "1", "krylov_rel_tol": "0.01"}
t0 = time.perf_counter()
solver = NewtonSolver(it, pd_info, workdir=os.path.join(tmp, "work"), dump=dump)
torch.cuda.synchronize(); t1 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
n = 0
while not solver.converged_flat() and n < 3:
    solver.step(); n += 1
torch.cuda.synchronize(); pr.disable(); t2 = time.perf_counter()
print(f"grid {nz}x{ny} dump={dump}: init {t1-t0:.2f} s, {n} Newton steps {t2-t1:.2f} s, Krylov iterations {[r.get('krylov_iterations') for r in solver.history[1:]]}, "
      f"|F|/|x| {[float((r['fcn_norm']/r['iterate_norm']).max()) for r in solver.history]}")
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
