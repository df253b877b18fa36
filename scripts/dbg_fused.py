"""debug driver: one small fused evaluation, compares with the stage-per-launch path"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "newton-krylov_ooc_b200")]
import numpy as np, torch
from oracle import nk_oracle as o
from nk_ooc_b200.py_driver_2d import modules
from nk_ooc_b200.spatial_axis import SpatialAxis
from nk_ooc_b200.engine import padded_members

nz, ny, B, nsteps = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (21, 33, 20, 3))]
ze = o.stretched_edges(nz, 0.0, 4000.0, 19.0); ye = o.stretched_edges(ny, 0.0, 50.0e5, 1.0)
tr = modules.Transport2D(SpatialAxis("depth", ze), SpatialAxis("ypos", ye))
m = modules.iage_model(tr)
m.set_uniform_schedule(nsteps)
rng = np.random.default_rng(0)
x = np.abs(rng.normal(size=(2, nz, ny, B)))
xd = torch.zeros((2, nz, ny, padded_members(B)), dtype=torch.float64, device="cuda")
xd[..., :B] = torch.from_numpy(x).cuda()
os.environ["NKB_FUSED"] = "0"
ref = m.eval(xd, B).cpu().numpy()[..., :B]
os.environ["NKB_FUSED"] = "1"
got = m.eval(xd, B).cpu().numpy()[..., :B]
torch.cuda.synchronize()
d = np.abs(got - ref)
print("dbg", os.environ.get("NKB_FUSED_DBG"), "OK max|diff| %.3e of %.3e" % (d.max(), np.abs(ref).max()), "argmax", np.unravel_index(d.argmax(), d.shape))
