import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "newton-krylov_ooc_b200")
import numpy as np, torch
from oracle import imex_oracle as im, nk_oracle as o
from nk_ooc_b200.py_driver_2d import modules
from nk_ooc_b200.spatial_axis import SpatialAxis
from nk_ooc_b200.engine import padded_members
nz, ny = 10, 7
ze = o.stretched_edges(nz, 0.0, 4000.0, 19.0); ye = o.stretched_edges(ny, 0.0, 50.0e5, 1.0)
g = o.Grid2D(ze, ye); tr = modules.Transport2D(SpatialAxis("depth", ze), SpatialAxis("ypos", ye))
rng = np.random.default_rng(11)
for B in [70]:
    mod, m = im.Module2D("phosphorus", g, phos=o.Phosphorus2D(g)), modules.phosphorus_model(tr)
    x = np.abs(rng.normal(size=(3, nz, ny, B))) * 0.5
    for nsteps in [1, 2]:
        m.set_uniform_schedule(nsteps)
        xd = torch.zeros((3, nz, ny, padded_members(B)), dtype=torch.float64, device="cuda"); xd[..., :B] = torch.from_numpy(x).cuda()
        got = m.eval(xd, B).cpu().numpy()[..., :B]
        want = im.model_year_2d(mod, x, nsteps)
        bad = np.argwhere(np.abs(got - want) > 1e-9 * np.abs(want).max())
        print("nsteps", nsteps, "bad count", len(bad)); print(bad[:50].T)
