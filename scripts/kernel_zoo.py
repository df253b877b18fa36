"""Launch every kernel of libnkb200.so once or twice at a production size, for a per-kernel ncu table
(profiles/r01_kernel_hbm_table.txt):

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/zoo.csv python scripts/kernel_zoo.py
    python scripts/kernel_zoo_table.py gpurun_out/zoo.csv > profiles/r01_kernel_hbm_table.txt

Without ncu it prints CUDA-event times of the same calls (python scripts/kernel_zoo.py --time)."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "newton-krylov_ooc_b200")]
import numpy as np
import torch

import bench
from nk_ooc_b200 import engine

TIME = "--time" in sys.argv
B = 4096


class A:
    pass


def timed(label, fn, alg_bytes=None, reps=3):
    fn()
    torch.cuda.synchronize()
    if not TIME:
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = f"{alg_bytes / ms / 1e6:8.0f} GB/s (algorithmic)" if alg_bytes else ""
    print(f"{label:58s} {ms:9.3f} ms {gbs}", flush=True)


def model_for(grid, module, nsteps):
    a = A()
    a.grid, a.module, a.nsteps = grid, module, nsteps
    return bench.build_model(a)


# ---- K1+K2: fused step (forced, iage), stage-per-launch TMA (phosphorus), plain stage kernel (B = 4) ----
for grid, module, nsteps, b in (("refined125x150", "forced", 12, B), ("default40x50", "iage", 24, B),
                                ("refined125x150", "phosphorus", 6, 2048), ("refined125x150", "forced", 6, 4)):
    model, depth, ypos = model_for(grid, module, nsteps)
    x = torch.rand(model.state_shape(b), dtype=torch.float64, device="cuda") + 0.5
    f = torch.empty_like(x)
    N = model.n
    timed(f"model_eval {grid} {module} B={b} ({nsteps} steps)", lambda: model.eval(x, b, out=f),
          alg_bytes=8.0 * N * b * (4 * nsteps + 1), reps=2)
    if module == "forced" and b == B:
        # ---- tendency (K1 alone), K5 wdot, K6 axpby, limiter, pack/unpack at the bench size ----
        timed("model_tend", lambda: model.tend(0.3 * model.t1, x, b), alg_bytes=16.0 * N * b)
        wgt = np.outer(depth.delta, ypos.delta)
        rw = engine.RegionWeights(np.ones((len(depth), len(ypos)), dtype=np.int32), wgt)
        y = torch.rand_like(x)
        timed("wdot (dot_prod)", lambda: rw.dot(x, y, b), alg_bytes=16.0 * N * b)
        timed("wdot (mean)", lambda: rw.dot(x, None, b), alg_bytes=8.0 * N * b)
        al = torch.rand((1, b), dtype=torch.float64, device="cuda")
        timed("axpby (per-region scalars)", lambda: rw.axpby(al, x, al, y, b), alg_bytes=24.0 * N * b)
        timed("limiter_scalef", lambda: rw.limiter_scalef(x, y, 0.0, None, b), alg_bytes=16.0 * N * b)
        maj = torch.rand((b, 1, len(depth), len(ypos)), dtype=torch.float64, device="cuda")
        timed("pack_members", lambda: engine.pack(maj), alg_bytes=16.0 * N * b)
        timed("unpack_members", lambda: engine.unpack(x, b), alg_bytes=16.0 * N * b)
        # ---- K4: (i) the per-column tridiagonal systems of a grid without lateral processes as ONE band
        # (150 independent blocks of nz rows), (ii) the same band fully coupled (one sequential chain),
        # (iii) the 2-D iage preconditioner's shape on the mid grid with ONE right-hand side ----
        n = len(depth) * len(ypos)
        ab = np.zeros((3, n))
        ab[1] = 2.5
        ab[0, 1:] = -1.0
        ab[2, :-1] = -1.0
        rhs = x.reshape(n, -1)
        fac = engine.BandedFactor(ab, 1, 1)
        timed("banded_solve kl=ku=1, one chain of nz*ny rows", lambda: fac.solve(rhs, b, scale=1.0 / model.t1, subtract_rhs=True),
              alg_bytes=16.0 * N * b)
        edge = np.arange(len(depth), n, len(depth))
        ab[0, edge] = 0.0
        ab[2, edge - 1] = 0.0
        fac = engine.BandedFactor(ab, 1, 1)
        timed(f"banded_solve kl=ku=1, {fac.n_blocks} column blocks", lambda: fac.solve(rhs, b, scale=1.0 / model.t1, subtract_rhs=True),
              alg_bytes=16.0 * N * b)
        n1, k1 = 80 * 100, 300
        ab = np.random.default_rng(0).normal(size=(2 * k1 + 1, n1))
        ab[k1] += 40.0
        fac = engine.BandedFactor(ab, k1, k1)
        r1 = torch.rand((n1, 1), dtype=torch.float64, device="cuda")
        timed(f"banded_solve n={n1} kl=ku={k1} B=1 (factor {8e-6 * n1 * (3 * k1 + 2):.0f} MB)", lambda: fac.solve(r1, 1),
              alg_bytes=8.0 * n1 * (3 * k1 + 2))
        r8 = torch.rand((n1, 32), dtype=torch.float64, device="cuda")
        timed(f"banded_solve n={n1} kl=ku={k1} B=32", lambda: fac.solve(r8, 32), alg_bytes=8.0 * n1 * (3 * k1 + 2))
        # ---- round 2: fused Gram-Schmidt / lin_comb over k = 8 basis vectors (single states: the Krylov case, and
        # a 64-member batch), the panel substitution kernel on the refined grid's 2-D preconditioner shape ----
        for bb in (1, 64):
            xb = torch.rand(model.state_shape(bb), dtype=torch.float64, device="cuda")
            basis = [torch.rand_like(xb) for _ in range(8)]
            wv = torch.rand_like(xb)
            co = torch.rand((8, 1, bb), dtype=torch.float64, device="cuda")
            ldbb = xb.shape[-1]
            timed(f"mgs k=8 B={bb} (w in registers, one cooperative launch)", lambda: rw.mgs(wv, basis, bb),
                  alg_bytes=8.0 * N * ldbb * 10)
            timed(f"lin_comb k=8 B={bb}", lambda: rw.lin_comb(co, basis, bb), alg_bytes=8.0 * N * ldbb * 9)
            del xb, basis, wv
        n2, k2 = 18750, 450
        ab = np.random.default_rng(1).normal(size=(2 * k2 + 1, n2)) * 0.01
        ab[k2] = 1.0 + np.abs(ab).sum(axis=0)
        fac = engine.BandedFactor(ab, k2, k2)
        for bb in (1, 32):
            rb = torch.rand((n2, engine.padded_members(bb)), dtype=torch.float64, device="cuda")
            timed(f"banded_solve (panel) n={n2} kl=ku={k2} B={bb} (factor columns {16e-6 * n2 * k2:.0f} MB)",
                  lambda: fac.solve(rb, bb), alg_bytes=16.0 * n2 * k2)
        del fac, rhs, maj, y
    del model, x, f
    torch.cuda.empty_cache()

# ---- test_problem: persistent column-year kernel ----
from nk_ooc_b200.spatial_axis import spatial_axis_from_defn
from nk_ooc_b200.test_problem.model_state import ModelState, gen_depth_axis_file

for names in ("iage", "phosphorus"):
    tmp = tempfile.mkdtemp()
    info = {"model_name": "test_problem", "tracer_module_names": names, "po4_s_restoring_opt": "1",
            "grid_vars_fname": os.path.join(tmp, "depth_axis.nc"), "depth_axisname": "depth", "reinvoke": "False"}
    gen_depth_axis_file(info, spatial_axis_from_defn("depth", nlevs=20))
    ModelState.configure(info)
    xs = ModelState.from_members([ModelState("gen_init_iterate")] * 1024)
    timed(f"test_problem {names} column year B=1024", lambda: xs.comp_fcn(None, None), reps=1)
    ModelState.reset()
