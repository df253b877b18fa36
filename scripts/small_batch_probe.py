"""per-step time of tiny batches (B = 1, 2, 4): fused step kernel (NKB_FUSED_MIN_B=1, default) against the
stage-per-launch kernels (NKB_FUSED_MIN_B=8)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "newton-krylov_ooc_b200")]
import torch, bench
class A: pass
for grid, module in (("refined125x150", "forced"), ("refined125x150", "iage"), ("refined125x150", "phosphorus"), ("ci30x30", "iage"), ("default40x50", "iage")):
    a = A(); a.grid, a.module, a.nsteps = grid, module, 240
    model, depth, ypos = bench.build_model(a)
    for b in (1, 2, 4):
        x = torch.rand(model.state_shape(b), dtype=torch.float64, device="cuda") + 0.5
        x[..., b:] = 0.0
        f = torch.empty_like(x)
        for minb in ("8", "1"):
            os.environ["NKB_FUSED_MIN_B"] = minb
            model.eval(x, b, out=f); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); model.eval(x, b, out=f); e1.record(); torch.cuda.synchronize()
            print(f"{grid} {module} B={b} MIN_B={minb}: {e0.elapsed_time(e1)/240*1e3:.1f} us per step, checksum {float(f[..., :b].abs().mean()):.6e}", flush=True)
