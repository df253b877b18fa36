#!/usr/bin/env python
"""bench.py — model-year evaluations per second of the batched F(x) = x(T) - x(0) hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--grid refined125x150|mid80x100|default40x50|ci30x30] [--module forced|iage|phosphorus]
                    [--members B_per_gpu] [--nsteps S]

A "step" is one pass of the hot path over one batch: B members integrated over one model
year (S time steps of the 2-stage IMEX scheme = S fused step launches).  Prints ONE JSON
line (rank 0).  Default workload: BASELINE.json configs[4] — py_driver_2d forced_o2_like on the
refined 125 x 150 synthetic grid, 4096 perturbed members, sharded over the GPUs (strong scaling: 512 per GPU
at 8 GPUs, with the all-gather of the result columns and the all-reduce of the residual norms inside the timed
region); --scaling weak keeps 4096 members per GPU with no collective.

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
    ap.add_argument("--members", type=int, default=4096, help="members in total (--scaling strong) or per GPU (weak)")
    ap.add_argument("--nsteps", type=int, default=0,
                    help="uniform time steps per model year; 0 = the production graded schedule (2640 steps)")
    ap.add_argument("--cpu-sample-seconds", type=float, default=0.0, help="(ignored: the CPU leg takes fixed-work samples)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the phosphorus block of the default run")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default; SURVEY.md 8d: 4096 members = 512 per GPU at 8 GPUs): --members in total, "
                         "sharded over the ranks, the gather of the result columns and the all-reduce of the residual "
                         "norms inside the timed region; weak: --members per GPU, no collective.  Identical at 1 GPU")
    return ap.parse_args()


def default_steps_per_year(args):
    """time steps per model year of the workload: --nsteps, else the product's graded schedule
    (nk_ooc_b200/py_driver_2d/model_state.py:default_schedule — 2640, 5280 for iage on grids finer than 60 levels)"""
    if args.nsteps:
        return int(args.nsteps)
    nz = GRIDS[args.grid][0]
    return 5280 if (args.module == "iage" and nz > 60) else 2640


def workload_config(args, world, S=None):
    """the `config` object of the JSON line — ONE function for both arms, so that the reference arm reports exactly
    the workload of ours (rank 0's shard under strong scaling)"""
    nz, ny, _ = GRIDS[args.grid]
    T = TRACERS[args.module]
    N = T * nz * ny
    strong = args.scaling == "strong" and world > 1
    if strong:
        B_total = args.members
        width = ((B_total + world - 1) // world + 31) // 32 * 32  # distributed.member_block_width
        B = min(B_total, width)
        per_gpu = f"{B} of {B_total} members per GPU (strong scaling)"
    else:
        B_total, B = world * args.members, args.members
        per_gpu = f"{B} perturbed members per GPU"
    return {
        "workload": f"py_driver_2d {args.module}{'_o2_like' if args.module == 'forced' else ''} on {args.grid} "
                    f"({nz}x{ny}, T={T}), {per_gpu}, one model year per step",
        "members_per_gpu": B, "members_total": B_total, "grid": args.grid, "module": args.module, "N": N,
        "time_steps_per_year": int(S if S is not None else default_steps_per_year(args)),
        "implicit_stages_per_step": 2, "scheme": "IMEX ARS(2,2,2)",
        "cache": "state batch (%.0f MB) larger than L2; no flush needed" % (8e-6 * N * B),
        "parallelism": (f"{B_total} members sharded over {world} GPU(s) in 32-aligned blocks; per step one NCCL "
                        f"all-gather of the result columns and one all-reduce of the residual norms, both inside "
                        f"the timed region" if strong else
                        f"members sharded over {world} GPU(s), no data-path collective"),
    }


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


CPU_SAMPLE_STEPS = 6     # accepted Radau steps per sample
CPU_SAMPLE_REPEATS = 3   # samples per worker; the median is used
CPU_RESTART_FRAC = 0.5   # the samples restart the integration at mid-year
CPU_LEG_TIMEOUT_S = float(os.environ.get("NKB_CPU_LEG_TIMEOUT_S", "900"))  # a CPU leg that has not delivered its samples by then is reported as unavailable
CPU_MAX_SPREAD = 1.2     # largest / smallest estimate among the samples of a worker before it is called unstable


def usable_cpus():
    """CPUs this process may run on: the affinity mask, cut to the cgroup CPU quota if there is one (a
    container with a 16-CPU quota on a 32-CPU host reports 32 from os.cpu_count(); 32 busy processes would
    then time-share and every per-core cost would double)"""
    cpus = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else list(range(os.cpu_count() or 1))
    quota = None
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:  # cgroup v2
            q, per = f.read().split()
            if q != "max":
                quota = float(q) / float(per)
    except (OSError, ValueError):
        try:
            with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f, open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as g:
                q, per = float(f.read()), float(g.read())
                if q > 0:
                    quota = q / per
        except (OSError, ValueError):
            pass
    if quota is not None and quota >= 1.0:
        cpus = cpus[: max(1, int(quota))]
    return cpus


def cpu_model_name():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _cpu_sample_worker(payload):
    """One member on one pinned CPU: CPU_SAMPLE_REPEATS samples of exactly CPU_SAMPLE_STEPS accepted steps of the
    reference's Radau stepper (instrumented, oracle/cpu_ref_profile.py) restarted at mid-year from the
    workload's state with a fixed first step (T/1000, the mean step of a full run) — a fixed amount of work
    instead of a wall-clock budget from t = 0, where the start-up transient of the step-size controller
    decided how many (and how small) steps a sample saw.  Returns the per-operation mean costs of every sample."""
    args_d, seed, cpu = payload
    os.environ["OMP_NUM_THREADS"] = "1"
    if cpu is not None and hasattr(os, "sched_setaffinity"):
        try:
            os.sched_setaffinity(0, {cpu})
        except OSError:
            pass
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
    samples = []
    for _ in range(CPU_SAMPLE_REPEATS):
        inst = Instrumented(mod, sparsity, x0, CPU_RESTART_FRAC * YEAR, YEAR, first_step=YEAR / 1000.0)
        wall = inst.run_steps(CPU_SAMPLE_STEPS)
        cost = {k: inst.times[k] / max(1, inst.counts[k]) for k in inst.times}
        # untimed remainder of the stepper (forming mu/h I - J before each factorisation, norms, step control):
        # charged per LU factorisation — measured proportional to their number (0.015-0.020 s per LU for 6, 12
        # and 24-step samples on 40 x 50 while the remainder per STEP varied 0.031-0.056 s with the share of
        # rejected steps in the sample)
        other = max(0.0, wall - sum(inst.times.values())) / max(1, inst.counts["lu"])
        frac = (inst.solver.t - CPU_RESTART_FRAC * YEAR) / YEAR
        samples.append({"cost": cost, "other": other, "counts": dict(inst.counts), "wall": wall, "frac": frac})
    return samples


def _estimate(sample, counts):
    """seconds for one full model-year evaluation from a sample's per-operation costs"""
    if counts is not None:
        c = counts["counts"]
        return sum(c[k] * sample["cost"][k] for k in sample["cost"]) + c["lu"] * sample["other"]
    return sample["wall"] / max(sample["frac"], 1e-12)  # no committed counts: linear in the sampled part of the year


def cpu_baseline(args, cpus):
    """Reference CPU algorithm (oracle port of py_driver_2d/model_state.py:102-114: scipy Radau, rtol = atol =
    1e-6, max_step = T/100, analytic sparse Jacobian + SuperLU) on a bounded, FIXED-WORK sample: one process
    per usable CPU (pinned), one member each; every process takes CPU_SAMPLE_REPEATS samples of
    CPU_SAMPLE_STEPS accepted Radau steps restarted at mid-year and keeps the median.  A full evaluation's time
    is the sampled per-operation costs (RHS, Jacobian, sparse LU, triangular solve, stepper overhead per LU)
    times the operation counts of one complete model-year run of the same workload (measured once in the
    build container, profiles/cpu_ref_counts.json).  evals/s = processes / median over processes.  A worker
    whose samples differ by more than 20 % is reported loudly (stderr + "unstable" in the JSON)."""
    import multiprocessing as mp

    args_d = {k: getattr(args, k) for k in ("grid", "module", "members", "nsteps")}
    procs = len(cpus)
    # always in worker processes: a Radau step cannot be interrupted from inside, and one step of the coupled
    # phosphorus system on refined125x150 (56 250 unknowns, dense column blocks) takes more than 15 minutes on
    # one core — such a workload is reported as unavailable instead of holding the bench for hours
    with mp.get_context("spawn").Pool(procs) as pool:
        pending = pool.map_async(_cpu_sample_worker, [(args_d, 1, cpu) for cpu in cpus])
        try:
            res = pending.get(timeout=CPU_LEG_TIMEOUT_S)
        except mp.TimeoutError:
            pool.terminate()
            why = (f"{CPU_SAMPLE_REPEATS} x {CPU_SAMPLE_STEPS} Radau steps of {args.grid}/{args.module} did not finish "
                   f"within {CPU_LEG_TIMEOUT_S:.0f} s on this host")
            sys.stderr.write(f"bench.py: CPU reference leg unavailable: {why}\n")
            return {"value": None, "unit": "model-year evals/s", "cores": procs, "kind": "port", "sample": why,
                    "unavailable": why, "cpu_model": cpu_model_name()}
    key = f"{args.grid}/{args.module}"
    cpath = os.path.join(ROOT, "profiles", "cpu_ref_counts.json")
    counts = None
    if os.path.exists(cpath):
        with open(cpath) as f:
            counts = json.load(f).get(key)
    if counts is not None and not counts.get("finished"):
        counts = None
    per_worker, spreads = [], []
    for samples in res:
        ests = sorted(_estimate(smp, counts) for smp in samples)
        per_worker.append(ests[len(ests) // 2])
        spreads.append(ests[-1] / ests[0])
    est = float(np.median(per_worker))
    unstable = procs == 1 and max(spreads) > CPU_MAX_SPREAD
    if unstable:
        sys.stderr.write(f"bench.py: CPU reference samples are NOT reproducible on this host: the estimates of the "
                         f"worker differ by a factor {max(spreads):.2f} (> {CPU_MAX_SPREAD})\n")
    s0 = res[0][len(res[0]) // 2]
    how = (f"x operation counts of one full model-year run of this workload {counts['counts']} "
           f"(build container, 1 core: {counts['wall_s_build_container_1core']:.0f} s measured there)"
           if counts is not None
           else "extrapolated linearly from the sampled fraction of the year (no committed counts for this workload)")
    sample = (f"{procs} member(s), one per pinned CPU; per member the median of {CPU_SAMPLE_REPEATS} samples of "
              f"{CPU_SAMPLE_STEPS} accepted scipy-Radau steps restarted at {CPU_RESTART_FRAC} yr (first step T/1000) "
              f"to measure per-operation costs on this host (rhs {s0['cost']['fun'] * 1e3:.2f} ms, "
              f"jac {s0['cost']['jac'] * 1e3:.2f} ms, LU {s0['cost']['lu'] * 1e3:.2f} ms, "
              f"solve {s0['cost']['solve'] * 1e3:.2f} ms, other {s0['other'] * 1e3:.2f} ms/LU) {how}; "
              f"{est:.0f} s per evaluation per core (median over processes; "
              f"min {min(per_worker):.0f}, max {max(per_worker):.0f}); sample spread within a worker <= {max(spreads):.2f}")
    return {"value": procs / est, "unit": "model-year evals/s", "cores": procs, "kind": "port", "sample": sample,
            "per_core_s_per_eval": est, "per_core_s_min_max": [float(min(per_worker)), float(max(per_worker))],
            "sample_spread_max": float(max(spreads)), "unstable": bool(unstable), "cpu_model": cpu_model_name()}


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


def _roofline(args, model, B, ms_per_step, launches, steps):
    """roofline object of one measured arm (B members per GPU, ms per model-year evaluation of the batch)"""
    N = model.T * model.nz * model.ny
    S, s = model.n_steps, 2
    bytes_alg_eval = 8.0 * N * (2 * s * S + 1)
    # launches per evaluation, counted by the library: the persistent fused step kernel integrates all S
    # time steps in ONE launch (+1 for the final difference); NKB_FUSED_PERSIST=0: S step launches;
    # stage-per-launch path: 2*S.  The roofline unit of work is one TIME STEP of the fused kernel
    # (one pass of the kernel's tile loop over the whole state batch) or one stage launch.
    launches_per_eval = max(1, int(round(launches / steps)))
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
    p3 = args.module == "phosphorus"
    return {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "peak_source": how,
        "kernel": ("nkb::step_fused_p3_kernel" if p3 else "nkb::step_fused_kernel") if fused else "nkb::stage_tma_kernel",
        "launches_per_eval": launches_per_eval,
        "unit_of_work": "one time step of the persistent fused step kernel (all members)" if fused
                        else "one stage launch",
        "alg_bytes_model": "8*N*(2*s*S+1) per member (SURVEY.md 8d: one read + one write of the state per "
                           "implicit stage); the fused step kernel moves less than that (see traffic)",
        "alg_bytes_per_launch": bytes_alg_eval * B / n_stage_launch, "avg_launch_ms": avg_launch_ms,
        "traffic_unit": "dram bytes per unit of work (ncu, profiles/traffic.json)",
        "limiter": ("only six consumer warps (tensor memory holds 3 x 125 levels for 64 (column, member) pairs per "
                    "SM), held by the four schedulers as 2 + 2 + 1 + 1: the FP64 issue slots of the two-warp schedulers "
                    "in the coupled source evaluations, then the shared-memory pipe at 80 % (issue slots 37 %, FP64 "
                    "pipe 28 %, DRAM 28 %; profiles/r02_ncu_full_step_fused_p3_refined125x150_B4096.txt, "
                    "profiles/r02_p3_variants.md)" if p3 else
                    "the board's 1000 W power cap (throughput follows the SM clock it leaves: ring-depth A/B in "
                    "profiles/r02_ncu_full_step_fused_persistent_refined125x150_forced_B4096.txt), then the "
                    "shared-memory LSU data pipe at 84 % of peak with every LDS at its ideal wavefront count; DRAM at "
                    "46 %: the fused kernel moves 0.57 of the algorithmic bytes") if fused
                    else "L2 round trip of the elimination intermediates",
    }


def _measure_device(model, x_dev, f_dev, B, steps, warmup, barrier, lib, after=None):
    """W untimed + K timed evaluations with inputs resident in HBM; CUDA events on the launching stream"""
    import torch

    for _ in range(warmup):
        model.eval(x_dev, B, out=f_dev)
        if after is not None:
            after()
    barrier()
    launches0 = lib.nkb_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        model.eval(x_dev, B, out=f_dev)
        if after is not None:
            after()
    ev1.record()
    barrier()
    return ev0.elapsed_time(ev1), lib.nkb_launch_count() - launches0


def run_ours(args):
    import torch
    import torch.distributed as dist

    from nk_ooc_b200 import _lib, distributed, engine

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
    strong = args.scaling == "strong" and world > 1  # (one GPU: the same workload either way, reported as weak)

    model, depth, ypos = build_model(args)
    T, nz, ny = model.T, model.nz, model.ny
    N = T * nz * ny
    if strong:
        # args.members members IN TOTAL, sharded in 32-aligned blocks (512 per GPU at 8 GPUs)
        B_total = args.members
        lo, hi = distributed.member_block_range(B_total, rank, world)
        B = hi - lo
        width = distributed.member_block_width(B_total, world)
        if B < 1:
            raise SystemExit("bench.py --scaling strong: fewer member blocks than ranks")
        all_members = members_host(initial_profile(args.module, depth, ypos), B_total, 1000)
        x_host = torch.from_numpy(np.ascontiguousarray(all_members[lo:hi])).pin_memory()
    else:
        # every rank owns its own B members (independent units; no data-path collective)
        B_total = world * args.members
        B = args.members
        width = engine.padded_members(B)
        x_host = torch.from_numpy(members_host(initial_profile(args.module, depth, ypos), B, 1000 + rank)).pin_memory()
    f_host = torch.empty_like(x_host).pin_memory()
    x_dev = torch.zeros((T, nz, ny, width), dtype=torch.float64, device="cuda")
    x_dev[..., : engine.padded_members(B)] = engine.pack(x_host.cuda())
    f_dev = torch.empty_like(x_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- strong scaling: the collectives of the Newton-Krylov driver inside the timed region ----
    coll = None
    after = None
    if strong and world > 1:
        weights = engine.RegionWeights(np.ones((nz, ny), dtype=np.int32), np.outer(depth.delta, ypos.delta))
        f_full = torch.zeros((T, nz, ny, engine.padded_members(B_total)), dtype=torch.float64, device="cuda")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        coll = {"gather_ms": [], "reduce_ms": []}
        pending = []

        def after():
            # (a) gather of the result columns to the owner of the Krylov basis: one all_gather_into_tensor of the
            # member-fastest blocks + nkb_interleave_blocks; (b) all-reduce (max over the ranks' members) of the
            # [n_modules, region_cnt] residual norms that the convergence test of the driver needs
            # the residual norms are reduced FIRST: the first collective after the model year also absorbs the
            # difference between the ranks' finishing times (boards differ in how hard their power cap bites),
            # which would otherwise be booked on the gather
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            evs[0].record()
            fl = f_dev.reshape(T, nz * ny, width)
            worst = torch.sqrt(weights.dot(fl, fl, B)).amax(dim=1).reshape(1, -1)
            dist.all_reduce(worst, op=dist.ReduceOp.MAX)
            evs[1].record()
            distributed.gather_member_blocks(f_dev, B_total, out=f_full)
            evs[2].record()
            pending.append(evs)

    # ---- device-resident arm ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches = _measure_device(model, x_dev, f_dev, B, args.steps, args.warmup, barrier, lib, after)
    clocks = sampler.stop() if rank == 0 else None
    if coll is not None:
        for evs in pending[args.warmup:]:
            coll["reduce_ms"].append(evs[0].elapsed_time(evs[1]))
            coll["gather_ms"].append(evs[1].elapsed_time(evs[2]))

    # ---- end-to-end arm: host buffers through the C ABI (H2D + pack + eval + unpack + D2H) ----
    e2e_steps = max(1, min(args.steps, 2))
    model.eval_host(x_host, f_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        model.eval_host(x_host, f_host)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps

    g_ms = float(np.median(coll["gather_ms"])) if coll else 0.0
    r_ms = float(np.median(coll["reduce_ms"])) if coll else 0.0
    tens = torch.tensor([ms, e2e_s, g_ms, r_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tens, op=dist.ReduceOp.MAX)
    ms, e2e_s, g_ms, r_ms = (float(v) for v in tens)
    checksum = float(f_host.double().abs().mean())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = ms / args.steps
    value = B_total / (ms_per_step * 1e-3)
    S, s = model.n_steps, 2
    # the roofline is the step kernel's: the collectives' share of the step is taken out of its time
    kernel_ms = ms_per_step - g_ms - r_ms
    per_gpu = f"{B} of {B_total} members per GPU (strong scaling)" if strong else f"{B} perturbed members per GPU"
    out = {
        "metric": "model-year evals/sec (batched perturbations)",
        "value": value,
        "unit": "model-year evals/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "strong" if strong else "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": dict(workload_config(args, world, S), members_per_gpu=B, members_total=B_total),
        "roofline": _roofline(args, model, B, kernel_ms, launches - (2 * args.steps if coll else 0), args.steps),
        "e2e": {
            "value": B_total / e2e_s, "unit": "model-year evals/s",
            "h2d_bytes_per_step": 8 * N * B_total, "d2h_bytes_per_step": 8 * N * B_total,
            "api": "nkb_model_eval_host (C ABI, pinned host buffers)",
        },
        "gpu_launches": int(launches),
        "clocks": clocks,
        "result_checksum": checksum,
    }
    if coll is not None:
        gbytes = 8.0 * N * width * (world - 1)  # received by every rank
        out["collectives"] = {
            "all_gather_ms": g_ms, "all_gather_bytes_received_per_rank": gbytes,
            "all_gather_GBps_per_rank": gbytes / (g_ms * 1e-3) / 1e9 if g_ms > 0 else None,
            "all_reduce_ms": r_ms, "all_reduce_bytes": 8,
            "all_reduce_includes": "nkb_wdot of the local residual norms and the wait for the slowest rank's model year",
            "share_of_step": (g_ms + r_ms) / ms_per_step,
            "limiting": "the all-gather (NVLink receive bandwidth of every rank: it gets the columns of all other "
                        "ranks); the 8-byte all-reduce is launch latency",
            "note": "max over ranks of the per-rank medians; all_gather includes nkb_interleave_blocks",
        }
    # ---- second workload of BASELINE.json configs[4] in the default run: the three coupled phosphorus tracers ----
    if world == 1 and not strong and args.module == "forced" and not args.no_extra:
        del model, x_dev, f_dev
        torch.cuda.empty_cache()
        out["extra"] = {"phosphorus": _extra_phosphorus(args, lib, barrier)}
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline(args, usable_cpus()[:1])
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def _extra_phosphorus(args, lib, barrier):
    """py_driver_2d phosphorus (T = 3) on the same grid and member count, device-resident and end to end;
    2 timed evaluations after 3 warm-ups (a model year of 4096 members takes ~5.4 s)"""
    import copy

    import torch

    from nk_ooc_b200 import engine

    a = copy.copy(args)
    a.module = "phosphorus"
    model, depth, ypos = build_model(a)
    B = a.members
    x_host = torch.from_numpy(members_host(initial_profile(a.module, depth, ypos), B, 2000)).pin_memory()
    f_host = torch.empty_like(x_host).pin_memory()
    x_dev = engine.pack(x_host.cuda())
    f_dev = torch.empty_like(x_dev)
    steps, warmup = 2, 3
    ms, launches = _measure_device(model, x_dev, f_dev, B, steps, warmup, barrier, lib)
    t0 = time.perf_counter()
    model.eval_host(x_host, f_host)
    e2e_s = time.perf_counter() - t0
    ms_per_step = ms / steps
    N = model.T * model.nz * model.ny
    return {
        "workload": f"py_driver_2d phosphorus on {a.grid} (T=3), {B} perturbed members, one model year per step",
        "value": B / (ms_per_step * 1e-3), "unit": "model-year evals/s", "ms_per_step": ms_per_step, "steps": steps,
        "warmup": warmup, "time_steps_per_year": model.n_steps,
        "e2e": {"value": B / e2e_s, "unit": "model-year evals/s", "h2d_bytes_per_step": 8 * N * B,
                "d2h_bytes_per_step": 8 * N * B},
        "roofline": _roofline(a, model, B, ms_per_step, launches, steps),
        "gpu_launches": int(launches),
        "result_checksum": float(f_host.double().abs().mean()),
    }


def run_reference(args):
    """the reference arm: the reference's own CPU algorithm on every usable host CPU.  A "step" is one
    fixed-work sample pass (see cpu_baseline); W untimed passes, then min(K, 5) passes whose median is the
    value (every pass is already a median of 3 samples per process)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cpus = usable_cpus()
    first = None
    for _ in range(min(args.warmup, 1) + 1):  # (the untimed pass, then the first of the timed ones)
        first = cpu_baseline(args, cpus)
        if first["value"] is None:
            print(json.dumps({"impl": "reference", "unavailable": first["unavailable"]}), flush=True)
            return
    vals = [first] + [cpu_baseline(args, cpus) for _ in range(max(1, min(args.steps, 5)) - 1)]
    vals.sort(key=lambda v: v["value"])
    cb = vals[len(vals) // 2]
    cb["passes"] = [v["value"] for v in vals]
    # reproducibility is judged on what is reported: the passes (each the median over the processes of their
    # median sample).  With every CPU busy the samples of ONE process scatter more (sample_spread_max: the
    # processes compete for memory bandwidth and cache), which the medians absorb.
    cb["pass_spread"] = vals[-1]["value"] / vals[0]["value"]
    cb["unstable"] = bool(cb["pass_spread"] > CPU_MAX_SPREAD)
    if cb["unstable"]:
        sys.stderr.write(f"bench.py --impl reference: the {len(vals)} passes differ by a factor {cb['pass_spread']:.2f} "
                         f"(> {CPU_MAX_SPREAD}): the CPU reference value is NOT reproducible on this host\n")
    nz, ny, _ = GRIDS[args.grid]
    T = TRACERS[args.module]
    out = {
        "impl": "reference",
        "metric": "model-year evals/sec (batched perturbations)",
        "value": cb["value"], "unit": "model-year evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
        "scaling": "strong" if (args.scaling == "strong" and world > 1) else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world),
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
