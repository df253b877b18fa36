"""time of the banded factorisation (set-up of a 2-D preconditioner): all SMs (NKB_BANDED_COOP=1, default for wide
blocks) against one CTA (NKB_BANDED_COOP=0)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "newton-krylov_ooc_b200")]
import numpy as np, torch
from nk_ooc_b200 import engine

for n, k in ((900, 90), (2000, 150), (8000, 300), (18750, 450)):
    rng = np.random.default_rng(0)
    ab = rng.normal(size=(2 * k + 1, n))
    ab[k] += 3.0 * np.sqrt(k)
    y = torch.rand((n, 1), dtype=torch.float64, device="cuda")
    for coop in ("1", "0"):
        if coop == "0" and n > 8000:
            continue
        os.environ["NKB_BANDED_COOP"] = coop
        torch.cuda.synchronize(); t0 = time.perf_counter()
        f = engine.BandedFactor(ab, k, k)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        x = f.solve(y, 1)
        print(f"n={n} kl=ku={k} coop={coop}: create {dt*1e3:9.1f} ms (incl. {ab.nbytes/1e6:.0f} MB upload), |x| {float(x.abs().max()):.6e}", flush=True)
        del f
