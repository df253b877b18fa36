"""Sharding of an ensemble over the GPUs of one box (one process per GPU, torch.distributed).

Members (perturbed states, Armijo candidates, coloured probes) are independent for the whole
model year, so the data path has NO collective: rank g owns the contiguous member block
[lo, hi).  Collectives carry only (a) the gather of result columns to every rank (or to the
owner of the Krylov basis) and (b) the all-reduce of [n_modules, region_cnt] partial dot
products when one state is split by module/region, and (c) the broadcast of a new iterate.
Backend: nccl over NVLink on GPUs, gloo in the CPU tests."""

import torch
import torch.distributed as dist


def member_range(n_members, rank, world):
    """contiguous, balanced block of members owned by `rank`: sizes differ by at most one"""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n_members, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_members(local, n_members, group=None):
    """all-gather member-major results: local [B_local, n] on every rank -> [n_members, n]"""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = member_range(n_members, rank, world)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local.shape[0]} members, expected {hi - lo}")
    width = max(member_range(n_members, r, world)[1] - member_range(n_members, r, world)[0] for r in range(world))
    padded = local.new_zeros((width,) + tuple(local.shape[1:]))
    padded[: hi - lo] = local
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    parts = []
    for r in range(world):
        rlo, rhi = member_range(n_members, r, world)
        parts.append(out[r][: rhi - rlo])
    return torch.cat(parts, dim=0)


def is_sharded(group=None):
    """True inside an initialised process group with more than one rank"""
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def _to_member_major(vals, B):
    """[..., ldb] member-fastest -> [B, n] (layout conversion for the gather: library kernel on the GPU)"""
    if vals.is_cuda:
        from . import engine

        return engine.unpack(vals, B).reshape(B, -1)
    return vals.movedim(-1, 0)[:B].reshape(B, -1).contiguous()


def _from_member_major(major, like):
    """[B, n] -> member-fastest tensor shaped like `like` ([..., ldb]); padding members are zero"""
    B = major.shape[0]
    if like.is_cuda:
        from . import engine

        fast = engine.pack(major.reshape((B,) + tuple(like.shape[:-1])))
        if fast.shape[-1] == like.shape[-1]:
            return fast
        out = torch.zeros_like(like)
        out[..., :B] = fast[..., :B]
        return out
    out = torch.zeros_like(like)
    out[..., :B] = major.reshape((B,) + tuple(like.shape[:-1])).movedim(0, -1)
    return out


def sharded_comp_fcn(state, group=None, hist_fname=None):
    """F(x) of a batched state (coloured probes, Armijo candidates, perturbed iterates) with its
    members sharded over the ranks: every rank holds the same `state`, evaluates the members of
    member_range(state.members, rank, world) — a model year each, no collective on the data path —
    and the result columns are all-gathered so that every rank returns the full batched F.
    Member results do not depend on which other members share a launch (tested bit for bit), so
    the sharded result equals the single-GPU one.  Outside a process group this is state.comp_fcn."""
    if not is_sharded(group):
        return state.comp_fcn(None, None, hist_fname)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B = state.members
    lo, hi = member_range(B, rank, world)
    res = state._like(clone_vals=False)
    local_fcn = None
    if hi > lo:
        local_fcn = state.member_slice(lo, hi).comp_fcn(None, None, hist_fname if lo == 0 else None)
    for ind, tms in enumerate(state.tracer_modules):
        n = tms.vals[..., 0].numel()
        if local_fcn is not None:
            local = _to_member_major(local_fcn.tracer_modules[ind].vals, hi - lo)
        else:
            local = tms.vals.new_zeros((0, n))
        full = gather_members(local, B, group)
        res.tracer_modules[ind].vals = _from_member_major(full, tms.vals)
    return res


def allreduce_sum(partial, group=None):
    """sum of [n_modules, region_cnt(, k)] partial dot products over the ranks (in place)"""
    dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


def broadcast_state(vals, src=0, group=None):
    """new iterate from the rank that owns it to everybody (in place)"""
    dist.broadcast(vals, src=src, group=group)
    return vals


def max_over_ranks(value, device, group=None):
    """device-timed durations are reported as the max over ranks"""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])
