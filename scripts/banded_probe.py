"""apply time of the wide-band solver (K4) on the refined grid's 2-D preconditioner shape: n = 18 750,
kl = ku = 450, B = 1 and 32 right-hand sides; panel kernel (default) against the row-by-row window kernel
(NKB_BANDED_PANEL=0).  python scripts/banded_probe.py [n kl]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "newton-krylov_ooc_b200"))
from nk_ooc_b200 import engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 18750
kl = ku = int(sys.argv[2]) if len(sys.argv) > 2 else 450
rng = np.random.default_rng(0)
ab = rng.normal(size=(kl + ku + 1, n)) * 0.01
ab[ku] = 1.0 + np.abs(ab).sum(axis=0)
for panel in ("1", "0"):
    os.environ["NKB_BANDED_PANEL"] = panel
    t0 = time.time()
    f = engine.BandedFactor(ab, kl, ku)
    torch.cuda.synchronize()
    t_fac = time.time() - t0
    for B in (1, 8, 32):
        y = torch.from_numpy(rng.normal(size=(n, engine.padded_members(B)))).cuda()
        x = f.solve(y, B)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(5):
            x = f.solve(y, B)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 5
        # residual of the solve against the band
        print(f"panel={panel} n={n} kl=ku={kl} B={B}: {ms:.3f} ms per solve (set-up {t_fac:.2f} s), "
              f"factor bytes {2 * n * kl * 8 / 1e6:.0f} MB -> {2 * n * kl * 8 / ms / 1e6:.1f} GB/s", flush=True)
    del f
