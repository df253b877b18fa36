"""CPU tests of the N>1 path with the gloo backend, world_size 2 (and 3): member sharding,
gather of result columns, all-reduce of partial dot products, broadcast of the iterate."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_members, tmpdir):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "newton-krylov_ooc_b200"))
    from nk_ooc_b200 import distributed as D

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = D.member_range(n_members, rank, world)
        n = 11
        full = torch.arange(n_members * n, dtype=torch.float64).reshape(n_members, n)
        # every rank "evaluates" its own members (here: a deterministic function of the member id)
        local = full[lo:hi] * 2.0 + 1.0
        gathered = D.gather_members(local, n_members)
        assert torch.equal(gathered, full * 2.0 + 1.0)
        # partial dot products of a state split over ranks
        part = torch.full((2, 3), float(rank + 1), dtype=torch.float64)
        D.allreduce_sum(part)
        assert torch.equal(part, torch.full((2, 3), float(sum(range(1, world + 1))), dtype=torch.float64))
        it = torch.full((5,), float(rank), dtype=torch.float64)
        D.broadcast_state(it, src=world - 1)
        assert torch.equal(it, torch.full((5,), float(world - 1), dtype=torch.float64))
        assert D.max_over_ranks(float(rank), "cpu") == float(world - 1)
        open(os.path.join(tmpdir, f"ok_{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_members", [(2, 7), (2, 8), (3, 4)])
def test_member_sharding_and_collectives_gloo(tmp_path, world, n_members):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_members, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"ok_{r}").exists()


def test_member_range_partitions_exactly():
    from nk_ooc_b200.distributed import member_range

    for n in (0, 1, 5, 4096, 4099):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = member_range(n, r, world)
                assert 0 <= lo <= hi <= n
                seen.extend(range(lo, hi))
            assert seen == list(range(n))
            sizes = [member_range(n, r, world)[1] - member_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        member_range(4, 2, 2)


def test_member_block_range_is_32_aligned_and_exact():
    """the partition of the data path: 32-aligned blocks (256-byte rows) so that the all-gathered blocks form
    one member-fastest batch; 4096 members over 8 ranks = 512 each"""
    from nk_ooc_b200.distributed import member_block_range, member_block_width

    assert member_block_width(4096, 8) == 512 and member_block_range(4096, 7, 8) == (3584, 4096)
    for n in (1, 5, 33, 70, 4096, 4099):
        for world in (1, 2, 3, 8):
            width = member_block_width(n, world)
            assert width % 32 == 0 and width * world >= n
            seen = []
            for r in range(world):
                lo, hi = member_block_range(n, r, world)
                assert lo == min(n, r * width) and hi - lo <= width
                seen.extend(range(lo, hi))
            assert seen == list(range(n))


# ---- sharded batched evaluation (distributed.sharded_comp_fcn) with a stand-in state ---------------
class _FakeTms:
    def __init__(self, vals):
        self.vals = vals


class _FakeState:
    """the surface sharded_comp_fcn uses: members, tracer_modules[i].vals [..., ldb], member_slice,
    comp_fcn, _like — F here is a deterministic per-member function so that any mix-up of members
    between ranks shows"""

    def __init__(self, vals_list, members):
        self.members = members
        self.tracer_modules = [_FakeTms(v) for v in vals_list]
        self.evaluated = 0

    def member_slice(self, lo, hi):
        out = []
        for tms in self.tracer_modules:
            v = torch.zeros(tms.vals.shape[:-1] + (hi - lo + 3,), dtype=torch.float64)  # own padding
            v[..., : hi - lo] = tms.vals[..., lo:hi]
            out.append(v)
        return _FakeState(out, hi - lo)

    def comp_fcn(self, res_fname, solver_state, hist_fname=None):
        self.evaluated += self.members
        return _FakeState([t.vals ** 2 + 3.0 * t.vals.sum(dim=tuple(range(t.vals.dim() - 1)), keepdim=True)
                           for t in self.tracer_modules], self.members)

    def _like(self, clone_vals=True):
        return _FakeState([t.vals.clone() if clone_vals else t.vals for t in self.tracer_modules], self.members)


def _worker_sharded(rank, world, port, n_members, tmpdir):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "newton-krylov_ooc_b200"))
    from nk_ooc_b200 import distributed as D

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gen = torch.Generator().manual_seed(5)
        ldb = n_members + 2
        vals = [torch.zeros((2, 4, 3, ldb), dtype=torch.float64), torch.zeros((1, 5, ldb), dtype=torch.float64)]
        for v in vals:
            v[..., :n_members] = torch.rand(v.shape[:-1] + (n_members,), generator=gen, dtype=torch.float64)
        state = _FakeState(vals, n_members)
        want = _FakeState([v.clone() for v in vals], n_members).comp_fcn(None, None)
        got = D.sharded_comp_fcn(state)
        for g, w in zip(got.tracer_modules, want.tracer_modules):
            assert torch.equal(g.vals[..., :n_members], w.vals[..., :n_members])
            assert torch.count_nonzero(g.vals[..., n_members:]) == 0
        open(os.path.join(tmpdir, f"ok_{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_members", [(2, 7), (2, 1), (3, 8), (2, 40), (3, 70), (2, 64)])
def test_sharded_comp_fcn_gloo(tmp_path, world, n_members):
    """every rank evaluates only its member block; the gathered result equals the unsharded one bit
    for bit, also when a rank owns no member (world 2, one member)"""
    port = _free_port()
    mp.spawn(_worker_sharded, args=(world, port, n_members, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"ok_{r}").exists()


def test_sharded_comp_fcn_without_process_group_is_plain_comp_fcn():
    from nk_ooc_b200 import distributed as D

    v = torch.rand((1, 3, 4), dtype=torch.float64)
    state = _FakeState([v], 4)
    got = D.sharded_comp_fcn(state)
    assert torch.equal(got.tracer_modules[0].vals, state.comp_fcn(None, None).tracer_modules[0].vals)
