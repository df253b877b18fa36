#!/usr/bin/env python
"""Error of the fixed-schedule integrator against the reference's Radau solution, per module and grid,
as a function of the number of time steps (GPU; writes a markdown table).

Truth where a golden exists (tests/golden/radau_<grid>_<module>.npz, the reference's own classes through its
own solve_ivp call at rtol = atol = 1e-9): 14x11, 30x30, 40x50 (+ 80x100 forced).  For grids without a Radau
truth (80x100, 125x150: hours of CPU per module) the same table is made against the scheme's own solution
on a 8x finer schedule (self-convergence), with the bench's synthetic states.

    python scripts/error_vs_steps.py > gpurun_out/error_vs_steps.md
"""
import glob
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "newton-krylov_ooc_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from nk_ooc_b200.engine import padded_members  # noqa: E402
from nk_ooc_b200.py_driver_2d import modules  # noqa: E402
from nk_ooc_b200.spatial_axis import SpatialAxis  # noqa: E402

RTOL, ATOL = 1.0e-3, 1.0e-6
SCHEDULES = [
    ("uniform 600", ("u", 600)), ("uniform 1200", ("u", 1200)), ("uniform 2400", ("u", 2400)),
    ("uniform 4800", ("u", 4800)), ("graded 1320 (10/60/120)", ("g", 10, 60, 120)),
    ("**graded 2640 (20/120/240, default)**", ("g", 20, 120, 240)), ("graded 5280 (40/240/480)", ("g", 40, 240, 480)),
]


def set_sched(model, spec):
    if spec[0] == "u":
        model.set_uniform_schedule(spec[1])
    else:
        model.set_graded_schedule(flat=spec[1], ramp=spec[2], ramp_first=spec[3])


def evaluate(model, x0):
    B = 2
    xd = torch.zeros(x0.shape + (padded_members(B),), dtype=torch.float64, device="cuda")
    xd[..., :B] = torch.from_numpy(np.ascontiguousarray(x0)).cuda()[..., None]
    f = model.eval(xd, B)
    torch.cuda.synchronize()
    model.check_health()
    return f[..., 0].cpu().numpy()


def ratio(got, want, x0):
    scale = np.maximum(1.0, np.abs(x0).reshape(x0.shape[0], -1).max(axis=1))[:, None, None]
    return float((np.abs(got - want) / (RTOL * np.abs(want) + ATOL * scale)).max())


def model_from_golden(g, module):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import radau_cases

    return radau_cases.model(g, module)


def table(title, rows, header):
    print(f"\n{title}\n")
    print("| " + " | ".join(header) + " |")
    print("|" + "---|" * len(header))
    for r in rows:
        print("| " + " | ".join(r) + " |")


def main():
    print("# error of F(x) vs number of time steps (max abs error; in brackets: worst ratio to the stated tolerance "
          "rtol 1e-3 |F| + atol 1e-6 max(1, max|x0|))")
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "radau_g*_*.npz")))
    cases = {}
    for f in files:
        _, grid, module = os.path.basename(f)[:-4].split("_")
        cases.setdefault(grid, {})[module] = np.load(f)
    for grid in sorted(cases, key=lambda s: int(s[1:].split("x")[0])):
        mods = [m for m in ("iage", "forced", "phosphorus") if m in cases[grid]]
        rows = []
        for label, spec in SCHEDULES:
            row = [label]
            for m in mods:
                g = cases[grid][m]
                truth = g["tol1e-09/fcn"] if "tol1e-09/fcn" in g.files else g["tol1e-06/fcn"]
                model = model_from_golden(g, m)
                set_sched(model, spec)
                got = evaluate(model, g["x0"])
                row.append(f"{np.abs(got - truth).max():.2e} ({ratio(got, truth, g['x0']):.3f})")
                del model
            rows.append(row)
        row = ["reference Radau at its own rtol = atol = 1e-6"]
        for m in mods:
            g = cases[grid][m]
            if "tol1e-09/fcn" in g.files and "tol1e-06/fcn" in g.files:
                row.append(f"{np.abs(g['tol1e-06/fcn'] - g['tol1e-09/fcn']).max():.2e} "
                           f"({ratio(g['tol1e-06/fcn'], g['tol1e-09/fcn'], g['x0']):.3f})")
            else:
                row.append("(truth is the 1e-6 run)")
        rows.append(row)
        row = ["max abs F"] + [f"{np.abs(cases[grid][m]['tol1e-09/fcn' if 'tol1e-09/fcn' in cases[grid][m].files else 'tol1e-06/fcn']).max():.3f}" for m in mods]
        rows.append(row)
        table(f"## {grid[1:]} against the reference's Radau solution (rtol = atol = 1e-9)", rows, ["schedule"] + mods)
    # self-convergence on the large grids (bench's synthetic workload)
    import bench

    for grid in ("mid80x100", "refined125x150"):
        rows = {label: [label] for label, _ in SCHEDULES}
        mods = ["iage", "forced", "phosphorus"]
        for m in mods:
            class A:
                pass

            a = A()
            a.grid, a.module, a.nsteps = grid, m, 0
            model, depth, ypos = bench.build_model(a)
            x0 = bench.members_host(bench.initial_profile(m, depth, ypos), 1, 5)[0]
            if m == "forced":  # an oxygen-minimum-like interior so that the sink limiter switches
                zz = depth.mid[:, None]
                yy = (ypos.mid / ypos.edges[-1])[None, :]
                omz = np.exp(-(((zz - 600.0) / 500.0) ** 2)) * (0.3 + 0.7 * np.sin(np.pi * yy) ** 2)
                x0 = np.maximum(x0 * (1.0 - 0.99 * omz), 0.004)
            model.set_graded_schedule(flat=160, ramp=960, ramp_first=1920)  # 21120 steps
            truth = evaluate(model, x0)
            for label, spec in SCHEDULES:
                set_sched(model, spec)
                got = evaluate(model, x0)
                rows[label].append(f"{np.abs(got - truth).max():.2e} ({ratio(got, truth, x0):.3f})")
            del model
        table(f"## {grid}: self-convergence against the same scheme on the 8x finer graded schedule (21120 steps)",
              list(rows.values()), ["schedule"] + mods)


if __name__ == "__main__":
    main()
