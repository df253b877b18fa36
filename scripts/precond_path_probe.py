import os, sys, tempfile
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, "newton-krylov_ooc_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from nk_ooc_b200.py_driver_2d.model_state import ModelState
from nk_ooc_b200.py_driver_2d.setup_solver import gen_grid_vars_file
from nk_ooc_b200.solver import ProbePreconditioner
from test_gpu_model_state import _modelinfo
for nz, ny, ratio in ((30, 30, "19.0"), (40, 50, "19.0"), (80, 100, "9.0")):
    tmp = tempfile.mkdtemp()
    info = _modelinfo(tmp, nz, ny); info["depth_delta_ratio_max"] = ratio
    gen_grid_vars_file(info); ModelState.configure(info)
    it = ModelState("gen_init_iterate")
    fcn = it.comp_fcn(None, None, os.path.join(tmp, "hist.nc"))
    it.gen_precond_jacobian(os.path.join(tmp, "hist.nc"), os.path.join(tmp, "precond.nc"))
    facs = ModelState._precond_factors(it.tracer_modules[0], os.path.join(tmp, "precond.nc"))
    print(nz, ny, "reference lateral preconditioner factors:", [f.path for f in facs], flush=True)
    if nz <= 40:
        pp = ProbePreconditioner(it, fcn)
        print(nz, ny, "probe preconditioner factor:", pp._factor.path, "n", pp._factor.n, flush=True)
    ModelState.reset()
