"""time the batched Thomas kernel on the per-column systems of a grid (env NKB_THOMAS_WPC / NKB_THOMAS_ZSMEM)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "newton-krylov_ooc_b200")]
import numpy as np, torch
from nk_ooc_b200 import engine

for nz, ny, B in ((125, 150, 4096), (40, 50, 4096), (20, 3, 4096)):
    n = nz * ny
    ab = np.zeros((3, n)); ab[1] = 2.5; ab[0, 1:] = -1.0; ab[2, :-1] = -1.0
    edge = np.arange(nz, n, nz); ab[0, edge] = 0.0; ab[2, edge - 1] = 0.0
    fac = engine.BandedFactor(ab, 1, 1)
    rhs = torch.rand((n, B), dtype=torch.float64, device="cuda")
    fac.solve(rhs, B, 0.5, True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = torch.empty_like(rhs)
    e0.record()
    for _ in range(5):
        engine.check(fac.lib.nkb_banded_solve(fac.handle, rhs.data_ptr(), out.data_ptr(), B, B, 0.5, 1, None), "solve")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{nz}x{ny} B={B} blocks={fac.n_blocks}: {ms:.3f} ms  {16.0*n*B/ms/1e6:.0f} GB/s", flush=True)
