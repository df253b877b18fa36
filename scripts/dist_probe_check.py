"""N-GPU check of the sharded batched evaluation (nk_ooc_b200/distributed.py) over NCCL:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/dist_probe_check.py

Every rank holds the same coloured-probe batch of the py_driver_2d column-regions grid (3 colours x 2
tracers x 20 levels = 120 members), evaluates its own member block for one model year and all-gathers
the result columns; the gathered F must equal the unsharded evaluation BIT FOR BIT on every rank
(members do not depend on which other members share a launch).  The probe preconditioner and a
speculative Armijo step are then built from sharded evaluations, the region-weighted dot products of a
state split by members are all-reduced, and a new iterate is broadcast."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "newton-krylov_ooc_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")

from nk_ooc_b200 import colouring, distributed as D
from nk_ooc_b200.py_driver_2d.model_state import ModelState
from nk_ooc_b200.py_driver_2d.setup_solver import gen_grid_vars_file
from nk_ooc_b200.solver import NewtonSolver, ProbePreconditioner
from test_gpu_model_state import _modelinfo

tmp = tempfile.mkdtemp(prefix=f"nkb_dist_{rank}_")
info = _modelinfo(tmp, 20, 3, "0.0", "0.0")
gen_grid_vars_file(info)
ModelState.configure(info)
iterate = ModelState("gen_init_iterate")
x0 = np.stack([iterate.get_tracer_vals("iage"), iterate.get_tracer_vals("iage_slow_rest")])
colour = colouring.column_colouring(3, 1)
probes = colouring.probe_batch(x0, colour, 1.0e-2)
B = probes.shape[0]
batched = ModelState("zeros", members=B)
batched.tracer_modules[0].vals[..., :B] = torch.from_numpy(np.ascontiguousarray(np.moveaxis(probes, 0, -1))).cuda()
whole = batched.comp_fcn(None, None).tracer_modules[0].vals[..., :B]
sharded = D.sharded_comp_fcn(batched).tracer_modules[0].vals[..., :B]
lo, hi = D.member_range(B, rank, world)
assert torch.equal(whole, sharded), f"rank {rank}: sharded evaluation differs from the unsharded one"
# dot products of a member-sharded state: partial sums over the own members, all-reduced
mine = batched.member_slice(lo, hi)
part = torch.from_numpy(np.asarray(mine.dot_prod(mine))).cuda().sum(dim=-1)  # [n_modules, R] summed over own members
D.allreduce_sum(part)
full = torch.from_numpy(np.asarray(batched.dot_prod(batched))).cuda().sum(dim=-1)
assert torch.allclose(part, full, rtol=1e-13, atol=0), (part, full)
# probe preconditioner and one Newton step with speculative Armijo: identical on every rank
pd_info = {"newton_rel_tol": "1.0e-5", "newton_max_iter": "5", "post_newton_fp_iter": "1", "krylov_rel_tol": "0.01"}
solver = NewtonSolver(iterate, pd_info, workdir=os.path.join(tmp, "work"), armijo_batch=3, dump=False,
                      precond_factory=lambda it, fcn: ProbePreconditioner(it, fcn))
solver.step()
vals = torch.from_numpy(np.stack([solver.iterate.get_tracer_vals("iage"),
                                  solver.iterate.get_tracer_vals("iage_slow_rest")])).cuda()
ref = vals.clone()
D.broadcast_state(ref, src=0)
assert torch.equal(vals, ref), f"rank {rank}: Newton iterate differs from rank 0"
fn = float(np.asarray(solver.fcn.norm()).max())
t = D.max_over_ranks(fn, "cuda")
if rank == 0:
    print(f"OK world={world}: {B} probes sharded {[D.member_range(B, r, world) for r in range(world)]}, "
          f"bit-identical gather, |F| after one probe-preconditioned Newton step {t:.3e}", flush=True)
dist.destroy_process_group()
