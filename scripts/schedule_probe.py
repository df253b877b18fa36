"""where the error of the graded schedule comes from (iage on the large grids): self-convergence against 21120 steps
for alternative (flat, ramp, ramp_first) step counts per hist interval.  python scripts/schedule_probe.py [grid]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "newton-krylov_ooc_b200"), os.path.join(ROOT, "scripts")]
import numpy as np
import bench
from error_vs_steps import evaluate, ratio

grid = sys.argv[1] if len(sys.argv) > 1 else "refined125x150"
module = sys.argv[2] if len(sys.argv) > 2 else "iage"
class A: pass
a = A(); a.grid, a.module, a.nsteps = grid, module, 0
model, depth, ypos = bench.build_model(a)
x0 = bench.members_host(bench.initial_profile(module, depth, ypos), 1, 5)[0]
model.set_graded_schedule(flat=160, ramp=960, ramp_first=1920)
truth = evaluate(model, x0)
SETS = ((20, 120, 240), (40, 240, 480), (20, 240, 480), (10, 240, 480), (40, 120, 240), (80, 120, 240),
        (20, 120, 480), (20, 120, 960), (20, 240, 240), (20, 480, 480), (30, 180, 360), (40, 160, 320))
if os.environ.get('SCHED_SETS') == 'coarse':
    SETS = ((20, 120, 240), (20, 60, 120), (20, 30, 60), (20, 20, 20), (10, 120, 240), (10, 60, 120), (30, 60, 120), (40, 40, 40))
for flat, ramp, first in SETS:
    model.set_graded_schedule(flat=flat, ramp=ramp, ramp_first=first)
    got = evaluate(model, x0)
    n = 48 * flat + 10 * ramp + 2 * first
    print(f"{grid} {module} flat {flat:3d} ramp {ramp:3d} first {first:3d}: {n:5d} steps, max abs err {np.abs(got - truth).max():.2e}, ratio {ratio(got, truth, x0):.3f}", flush=True)
