"""test_problem (1-D column) throughput probe: B members of iage / dye_decay / phosphorus through the
persistent column-year kernel (Richardson pair, production step counts)"""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "newton-krylov_ooc_b200")]
import numpy as np, torch
from nk_ooc_b200.spatial_axis import spatial_axis_from_defn
from nk_ooc_b200.test_problem.model_state import ModelState, gen_depth_axis_file

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for names in ("iage", "dye_decay_{suff}:010", "phosphorus"):
    tmp = tempfile.mkdtemp()
    info = {"model_name": "test_problem", "tracer_module_names": names, "po4_s_restoring_opt": "1",
            "grid_vars_fname": os.path.join(tmp, "depth_axis.nc"), "depth_axisname": "depth", "reinvoke": "False"}
    gen_depth_axis_file(info, spatial_axis_from_defn("depth", nlevs=20))
    ModelState.configure(info)
    x1 = ModelState("gen_init_iterate")
    x = ModelState.from_members([x1] * B)
    x.comp_fcn(None, None)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        f = x.comp_fcn(None, None)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"{names:24s} B={B}: {dt*1e3:8.1f} ms per batch, {B/dt:10.1f} model-year evals/s")
    ModelState.reset()
