#!/usr/bin/env python
"""bench.py — model-year evaluations per second of the batched F(x) = x(T) - x(0) hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--grid refined125x150|mid80x100|default40x50|ci30x30] [--module forced|iage|phosphorus]
                    [--members B_per_gpu] [--nsteps S]

A "step" is one pass of the hot path over one batch: B members integrated over one model
year (S time steps of the 2-stage IMEX scheme = S fused step launches).  Prints ONE JSON
line (rank 0).  Default workload: BASELINE.json configs[4] — py_driver_2d forced_o2_like on the
refined 125 x 150 synthetic grid, 4096 perturbed members per GPU (weak scaling).

--impl reference times the reference's own CPU algorithm (scipy solve_ivp Radau through
oracle/nk_oracle.py, the "port" of nk_ooc/py_driver_2d/model_state.py:102-114) on all host
cores on a bounded sample of the same workload.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "newton-krylov_ooc_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

YEAR = 365.0 * 86400.0
GRIDS = {  # nz, ny, depth delta_ratio_max (input/py_driver_2d/model_params.cfg:9-29)
    "refined125x150": (125, 150, 11.8),
    "mid80x100": (80, 100, 9.0),
    "default40x50": (40, 50, 19.0),
    "ci30x30": (30, 30, 19.0),
}
TRACERS = {"forced": 1, "iage": 2, "phosphorus": 3}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", default="refined125x150", choices=sorted(GRIDS))
    ap.add_argument("--module", default="forced", choices=sorted(TRACERS))
    ap.add_argument("--members", type=int, default=4096, help="members per GPU")
    ap.add_argument("--nsteps", type=int, default=0,
                    help="uniform time steps per model year; 0 = the production graded schedule (2640 steps)")
    ap.add_argument("--cpu-sample-seconds", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def axes(grid):
    from nk_ooc_b200.spatial_axis import SpatialAxis, edges_from_defn

    nz, ny, ratio = GRIDS[grid]
    depth = SpatialAxis("depth", edges_from_defn(nz, 0.0, 4000.0, ratio))
    ypos = SpatialAxis("ypos", edges_from_defn(ny, 0.0, 50.0e5, 1.0))
    return depth, ypos


def synthetic_forcing(depth, ypos, seed=4):
    """o2_like sink record shaped like input/py_driver_2d/po4_sms.nc ([61, nz, ny], 61 times over
    the year), synthetic: surface-intensified consumption with a seasonal cycle + seeded noise.
    scalef (-1/3) already applied (scripts/run_py_driver_2d_forced_o2_like.sh:14-25)."""
    rng = np.random.default_rng(seed)
    nt = 61
    times = np.linspace(0.0, YEAR, nt)
    prof = np.exp(-depth.mid / 400.0)[None, :, None]
    lat = np.exp(-(((ypos.mid - 2.5e6) / 1.5e6) ** 2))[None, None, :]
    season = (1.0 + 0.5 * np.sin(2 * np.pi * times / YEAR))[:, None, None]
    data = -(1.0 / 3.0) * 2.0e-8 * prof * lat * season * (1.0 + 0.1 * rng.random((nt, len(depth), len(ypos))))
    return times, np.ascontiguousarray(data)


def initial_profile(module, depth, ypos):
    """gen_init_iterate profiles of input/py_driver_2d/tracer_module_defs.yaml:7-48"""
    nz, ny = len(depth), len(ypos)
    if module == "forced":
        cols = [np.full(nz, 1.0)]
    elif module == "iage":
        cols = [np.interp(depth.mid, [55.0, 200.0], [0.0, 2.0])] * 2
    else:
        cols = [
            np.interp(depth.mid, [1.3e2, 2.6e2], [5.5e-3, 4.1]),
            np.interp(depth.mid, [9.5e1, 1.4e2], [7.1e-2, 1.5e-4]),
            np.interp(depth.mid, [1.7e2, 2.5e2], [1.8e-2, 7.9e-4]),
        ]
    return np.stack([np.broadcast_to(c[:, None], (nz, ny)) for c in cols]).astype(np.float64)


def members_host(x0, B, seed):
    """member-major [B, T, nz, ny]: x0 + sigma*v_b, sigma = 1e-4*||x0|| (model_state_base.py:509)"""
    rng = np.random.default_rng(seed)
    sigma = 1.0e-4 * np.sqrt(np.mean(x0 * x0)) or 1.0e-4
    out = np.empty((B,) + x0.shape)
    for b in range(B):
        out[b] = x0 + sigma * rng.standard_normal(x0.shape)
    return out


def build_model(args):
    from nk_ooc_b200.py_driver_2d import modules

    depth, ypos = axes(args.grid)
    tr = modules.Transport2D(depth, ypos, 0.1, 1000.0)
    if args.module == "forced":
        ft, fd = synthetic_forcing(depth, ypos)
        model = modules.forced_model(tr, "const", 1.0, 1.0 / 3600.0, "file", sms_times=ft, sms_data=fd, sink_thres=0.05)
    elif args.module == "iage":
        model = modules.iage_model(tr)
    else:
        model = modules.phosphorus_model(tr)
    if args.nsteps > 0:
        model.set_uniform_schedule(args.nsteps)
    else:
        model.set_graded_schedule()
    return model, depth, ypos


def oracle_module(args):
    from oracle import nk_oracle as o

    depth, ypos = axes(args.grid)
    g = o.Grid2D(depth.edges, ypos.edges, 0.1, 1000.0)
    if args.module == "forced":
        ft, fd = synthetic_forcing(depth, ypos)
        return o.Forced2D(g, restore_rate_10m=1.0 / 3600.0, restore_const=1.0, sms_opt="file", sms_times=ft,
                          sms_data=fd, sink_thres=0.05), depth, ypos
    if args.module == "iage":
        return o.Iage2D(g), depth, ypos
    return o.Phosphorus2D(g), depth, ypos


def _cpu_sample_worker(payload):
    """One member: run the reference's Radau stepper (instrumented, oracle/cpu_ref_profile.py) for
    about budget_s seconds at the workload's size and return the mean cost of each operation
    (RHS, Jacobian, sparse LU, triangular solve) plus the un-attributed Python overhead per step."""
    args_d, seed, budget_s = payload
    os.environ["OMP_NUM_THREADS"] = "1"
    from scipy import sparse

    from oracle.cpu_ref_profile import Instrumented

    class A:  # argparse-like
        pass

    a = A()
    a.__dict__.update(args_d)
    mod, depth, ypos = oracle_module(a)
    x0 = members_host(initial_profile(a.module, depth, ypos), 1, seed)[0].reshape(-1)
    r, c, _ = sparse.find(mod.comp_jacobian(0.0, x0))
    sparsity = sparse.csr_matrix((np.ones(r.shape), (r, c)))
    inst = Instrumented(mod, sparsity, x0, 0.0, YEAR)
    wall = inst.run(budget_s)
    cost = {k: inst.times[k] / max(1, inst.counts[k]) for k in inst.times}
    other = max(0.0, wall - sum(inst.times.values())) / max(1, inst.counts["step"])
    return cost, other, inst.counts, wall, inst.solver.t / YEAR


def cpu_baseline(args, procs, budget_s):
    """Reference CPU algorithm (oracle port of py_driver_2d/model_state.py:102-114: scipy Radau,
    rtol=atol=1e-6, max_step=T/100, analytic sparse Jacobian + SuperLU) on a bounded sample:
    `procs` processes, one member each, ~budget_s seconds of the real integration per process to
    measure the per-operation costs on THIS host; a full evaluation's wall time is those costs
    times the operation counts of one complete model-year run of the same workload (measured once
    in the build container, profiles/cpu_ref_counts.json).  evals/s = procs / mean(full wall)."""
    import multiprocessing as mp

    args_d = {k: getattr(args, k) for k in ("grid", "module", "members", "nsteps")}
    if procs == 1:
        res = [_cpu_sample_worker((args_d, 1, budget_s))]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_cpu_sample_worker, [(args_d, 1 + i, budget_s) for i in range(procs)])
    key = f"{args.grid}/{args.module}"
    cpath = os.path.join(ROOT, "profiles", "cpu_ref_counts.json")
    counts = None
    if os.path.exists(cpath):
        with open(cpath) as f:
            counts = json.load(f).get(key)
    ests = []
    for cost, other, cnt, wall, frac in res:
        if counts is not None and counts.get("finished"):
            c = counts["counts"]
            ests.append(sum(c[k] * cost[k] for k in cost) + c["step"] * other)
        else:  # no committed counts for this workload: linear extrapolation of the sampled progress
            ests.append(wall / max(frac, 1e-12))
    est = float(np.mean(ests))
    cost0, other0, cnt0, wall0, frac0 = res[0]
    how = (f"x operation counts of one full model-year run of this workload {counts['counts']} "
           f"(build container, 1 core: {counts['wall_s_build_container_1core']:.0f} s measured there)"
           if counts is not None and counts.get("finished")
           else "extrapolated linearly from the sampled fraction of the year (no committed counts for this workload)")
    sample = (f"{procs} member(s), one per process; ~{budget_s:.0f} s of the real scipy-Radau integration per member "
              f"({cnt0['step']} steps, {frac0:.2e} yr) to measure per-operation costs on this host "
              f"(rhs {cost0['fun'] * 1e3:.2f} ms, jac {cost0['jac'] * 1e3:.2f} ms, LU {cost0['lu'] * 1e3:.2f} ms, "
              f"solve {cost0['solve'] * 1e3:.2f} ms, other {other0 * 1e3:.2f} ms/step) {how}; "
              f"estimated {est:.0f} s per evaluation per core")
    return {"value": procs / est, "unit": "model-year evals/s", "cores": procs, "kind": "port", "sample": sample}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def run_ours(args):
    import torch
    import torch.distributed as dist

    from nk_ooc_b200 import _lib, engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    engine.require_cuda()
    torch.cuda.set_device(local)
    # stdout carries ONE JSON line: everything else a library writes to fd 1 (NCCL's version banner, ...)
    # goes to stderr; the JSON line is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()

    model, depth, ypos = build_model(args)
    B = args.members
    T, nz, ny = model.T, model.nz, model.ny
    N = T * nz * ny
    # every rank owns its own B members (independent units; no data-path collective)
    x_host = torch.from_numpy(members_host(initial_profile(args.module, depth, ypos), B, 1000 + rank)).pin_memory()
    f_host = torch.empty_like(x_host).pin_memory()
    x_dev = engine.pack(x_host.cuda())
    f_dev = torch.empty_like(x_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ----
    for _ in range(args.warmup):
        model.eval(x_dev, B, out=f_dev)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.nkb_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        model.eval(x_dev, B, out=f_dev)
    ev1.record()
    barrier()
    launches = lib.nkb_launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end arm: host buffers through the C ABI (H2D + pack + eval + unpack + D2H) ----
    e2e_steps = max(1, min(args.steps, 2))
    model.eval_host(x_host, f_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        model.eval_host(x_host, f_host)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps

    tens = torch.tensor([ms, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tens, op=dist.ReduceOp.MAX)
    ms, e2e_s = float(tens[0]), float(tens[1])
    checksum = float(f_host.double().abs().mean())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    S, s = model.n_steps, 2
    bytes_alg_eval = 8.0 * N * (2 * s * S + 1)
    # launches per evaluation, counted by the library: the persistent fused step kernel integrates all S
    # time steps in ONE launch (+1 for the final difference); NKB_FUSED_PERSIST=0: S step launches;
    # stage-per-launch path: 2*S.  The roofline unit of work is one TIME STEP of the fused kernel
    # (one pass of the kernel's tile loop over the whole state batch) or one stage launch.
    launches_per_eval = max(1, int(round(launches / args.steps)))
    fused = launches_per_eval < 2 * S  # (the phosphorus path adds two layout-conversion launches per evaluation)
    n_stage_launch = S if fused else 2 * S
    avg_launch_ms = ms_per_step / n_stage_launch
    peak, how = measured_peak_gbs()
    achieved = (bytes_alg_eval * B / n_stage_launch) / (avg_launch_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(f"{args.grid}/{args.module}/{B}")
    out = {
        "metric": "model-year evals/sec (batched perturbations)",
        "value": value,
        "unit": "model-year evals/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": f"py_driver_2d {args.module}{'_o2_like' if args.module == 'forced' else ''} on {args.grid} "
                        f"({nz}x{ny}, T={T}), {B} perturbed members per GPU, one model year per step",
            "members_per_gpu": B, "grid": args.grid, "module": args.module, "N": N,
            "time_steps_per_year": S, "implicit_stages_per_step": s, "scheme": "IMEX ARS(2,2,2)",
            "cache": "state batch (%.0f MB) larger than L2; no flush needed" % (8e-6 * N * B),
            "parallelism": f"members sharded over {world} GPU(s), no data-path collective",
        },
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": how,
            "kernel": ("nkb::step_fused_p3_kernel" if args.module == "phosphorus" else "nkb::step_fused_kernel")
                      if fused else "nkb::stage_tma_kernel",
            "launches_per_eval": launches_per_eval,
            "unit_of_work": "one time step of the persistent fused step kernel (all members)" if fused
                            else "one stage launch",
            "alg_bytes_model": "8*N*(2*s*S+1) per member (SURVEY.md 8d: one read + one write of the state per "
                               "implicit stage); the fused step kernel moves less than that (see traffic)",
            "alg_bytes_per_launch": bytes_alg_eval * B / n_stage_launch, "avg_launch_ms": avg_launch_ms,
            "traffic_unit": "dram bytes per unit of work (ncu, profiles/traffic.json)",
            "limiter": ("instruction latency of the six consumer warps (tensor memory holds 3 x 125 levels for only 64 "
                        "(column, member) pairs per SM): issue slots 37 %, FP64 pipe 30 %, DRAM 47 % "
                        "(profiles/r01_ncu_full_step_fused_p3_*.txt)" if args.module == "phosphorus" else
                        "shared-memory LSU data pipe at 84 % of peak (ncu l1tex__data_pipe_lsu_wavefronts), DRAM at 46 %: "
                        "the fused kernel moves 0.57 of the algorithmic bytes (profiles/r01_ncu_full_step_fused_persistent_*.txt)")
                       if fused else "L2 round trip of the elimination intermediates",
        },
        "e2e": {
            "value": world * B / e2e_s, "unit": "model-year evals/s",
            "h2d_bytes_per_step": 8 * N * B * world, "d2h_bytes_per_step": 8 * N * B * world,
            "api": "nkb_model_eval_host (C ABI, pinned host buffers)",
        },
        "gpu_launches": int(launches),
        "clocks": clocks,
        "result_checksum": checksum,
    }
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline(args, 1, args.cpu_sample_seconds)
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    vals = []
    for _ in range(max(1, min(args.steps, 2))):
        vals.append(cpu_baseline(args, procs, args.cpu_sample_seconds))
    cb = vals[-1]
    nz, ny, _ = GRIDS[args.grid]
    T = TRACERS[args.module]
    out = {
        "impl": "reference",
        "metric": "model-year evals/sec (batched perturbations)",
        "value": cb["value"], "unit": "model-year evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": f"py_driver_2d {args.module}{'_o2_like' if args.module == 'forced' else ''} on {args.grid} "
                        f"({nz}x{ny}, T={T}), {args.members} perturbed members per GPU, one model year per step",
            "members_per_gpu": args.members, "grid": args.grid, "module": args.module, "N": T * nz * ny,
        },
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "model-year evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
