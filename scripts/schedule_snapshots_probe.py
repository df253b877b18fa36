"""worst ratio to the stated tolerance of F AND of the hist snapshots against the reference's Radau golden, for
candidate graded schedules of iage on the 80 x 100 grid (tests/golden/radau_g80x100_iage.npz)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "newton-krylov_ooc_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import test_gpu_radau_parity as t

grid, module = (sys.argv[1], sys.argv[2]) if len(sys.argv) > 2 else ("g80x100", "iage")
g = t._load(os.path.join(ROOT, "tests", "golden"), grid, module)
x0, truth = g["x0"], g["tol1e-09/fcn"]
idx = [int(i) for i in g["tol1e-09/snap_idx"]]
SETS = ((20, 120, 240), (40, 120, 240), (40, 180, 360), (40, 240, 240), (40, 240, 480), (30, 240, 480), (40, 200, 400))
if os.environ.get('SCHED_SETS') == 'first':
    SETS = ((20, 120, 240), (20, 120, 120), (20, 120, 480), (20, 100, 100), (24, 120, 120))
for flat, ramp, first in SETS:
    model = t._model(g, module)
    model.set_graded_schedule(flat=flat, ramp=ramp, ramp_first=first)
    got, snaps = t._eval(model, x0, idx)
    rs = [t._tol_ratio(snaps[i], g["tol1e-09/snaps"][i], x0) for i in range(len(idx))]
    print(f"{grid} {module} {flat}/{ramp}/{first}: {48*flat+10*ramp+2*first} steps, F ratio {t._tol_ratio(got, truth, x0):.3f}, "
          f"worst snapshot ratio {max(rs):.3f} (snapshot {idx[int(np.argmax(rs))]})", flush=True)
    del model
