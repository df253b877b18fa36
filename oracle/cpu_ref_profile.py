"""TEST/BENCH INFRASTRUCTURE: operation counts of ONE full model-year evaluation with the
reference's CPU algorithm (scipy Radau exactly as nk_ooc/py_driver_2d/model_state.py:102-114
calls it), instrumented per operation.  Run once in the build container per benchmark
workload; the counts are committed in profiles/cpu_ref_counts.json and used by bench.py's
cpu_baseline leg, which measures the per-operation costs live on the GPU box's host cores.

    python -m oracle.cpu_ref_profile --grid refined125x150 --module forced
"""

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "newton-krylov_ooc_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

YEAR = 365.0 * 86400.0


class Instrumented:
    """scipy Radau stepper with per-operation counters/timers (fun, jac, lu, solve_lu)"""

    def __init__(self, mod, sparsity, y0, t_start, t_end, first_step=None):
        from scipy import integrate

        self.reset()

        def timed(name, f):
            def wrapper(*a, **k):
                t0 = time.perf_counter()
                out = f(*a, **k)
                self.times[name] += time.perf_counter() - t0
                self.counts[name] += 1
                return out

            return wrapper

        self.solver = integrate.Radau(
            timed("fun", mod.comp_tend), t_start, y0, t_end, max_step=YEAR * 0.01, rtol=1.0e-6, atol=1.0e-6,
            jac=timed("jac", mod.comp_jacobian), jac_sparsity=sparsity, first_step=first_step,
        )
        self.solver.lu = timed("lu", self.solver.lu)
        self.solver.solve_lu = timed("solve", self.solver.solve_lu)

    def reset(self):
        self.counts = {"fun": 0, "jac": 0, "lu": 0, "solve": 0, "step": 0}
        self.times = {"fun": 0.0, "jac": 0.0, "lu": 0.0, "solve": 0.0}

    def run_steps(self, n_steps):
        """exactly n_steps accepted Radau steps (or to the end of the interval); wall seconds"""
        t0 = time.perf_counter()
        for _ in range(n_steps):
            if self.solver.status != "running":
                break
            self.solver.step()
            self.counts["step"] += 1
        return time.perf_counter() - t0

    def run(self, budget_s=None):
        t0 = time.perf_counter()
        while self.solver.status == "running":
            if budget_s is not None and time.perf_counter() - t0 > budget_s:
                break
            self.solver.step()
            self.counts["step"] += 1
        return time.perf_counter() - t0


def main():
    import bench

    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="refined125x150")
    ap.add_argument("--module", default="forced")
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    args.members, args.nsteps = 1, 2400
    from scipy import sparse

    mod, depth, ypos = bench.oracle_module(args)
    x0 = bench.members_host(bench.initial_profile(args.module, depth, ypos), 1, args.seed)[0].reshape(-1)
    r, c, _ = sparse.find(mod.comp_jacobian(0.0, x0))
    sparsity = sparse.csr_matrix((np.ones(r.shape), (r, c)))
    inst = Instrumented(mod, sparsity, x0, 0.0, YEAR)
    wall = inst.run()
    rec = {
        "grid": args.grid, "module": args.module, "seed": args.seed, "wall_s_build_container_1core": wall,
        "counts": inst.counts, "times_s": inst.times,
        "nfev": inst.solver.nfev, "njev": inst.solver.njev, "nlu": inst.solver.nlu,
        "finished": inst.solver.status == "finished",
    }
    path = os.path.join(ROOT, "profiles", "cpu_ref_counts.json")
    data = {}
    if os.path.exists(path):
        with open(path) as f:
            data = json.load(f)
    data[f"{args.grid}/{args.module}"] = rec
    with open(path, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
