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


def member_block_width(n_members, world):
    """members per rank of the 32-aligned block partition used on the data path: rank g owns the members
    [g*W, min(B, (g+1)*W)), W a multiple of 32 (256-byte rows of the member-fastest layout), so that the
    all-gathered blocks line up as ONE member-fastest batch with members in their original order"""
    per = -(-n_members // world)
    return ((per + 31) // 32) * 32


def member_block_range(n_members, rank, world):
    width = member_block_width(n_members, world)
    lo = min(n_members, rank * width)
    return lo, min(n_members, lo + width)


def gather_member_blocks(local, n_members, group=None, out=None):
    """all-gather of member-fastest result blocks: `local` [..., W] on every rank (W =
    member_block_width; lanes beyond the rank's own members are ignored) -> [..., ldb] with all
    n_members in order.  One all_gather_into_tensor of the packed blocks and one library kernel
    (nkb_interleave_blocks) — no member-major staging, no per-rank Python lists."""
    world = dist.get_world_size(group)
    width = member_block_width(n_members, world)
    if local.shape[-1] != width:
        raise ValueError(f"local block has {local.shape[-1]} member lanes, expected {width}")
    lead = tuple(local.shape[:-1])
    n = 1
    for d in lead:
        n *= d
    gathered = torch.empty((world * n, width), dtype=local.dtype, device=local.device)  # [G][n][W], concatenated
    dist.all_gather_into_tensor(gathered, local.reshape(n, width).contiguous(), group=group)
    gathered = gathered.view(world, n, width)
    ldb = 1 if n_members == 1 else ((n_members + 31) // 32) * 32
    if out is None:
        out = torch.zeros(lead + (ldb,), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        from . import _lib, engine

        _lib.check(_lib.load().nkb_interleave_blocks(gathered.data_ptr(), out.data_ptr(), n, world, width, out.shape[-1],
                                                     n_members, engine._stream_ptr()), "nkb_interleave_blocks")
    else:  # gloo tests of the host logic
        full = gathered.permute(1, 0, 2).reshape(n, world * width)[:, :n_members]
        out.reshape(n, out.shape[-1])[:, :n_members] = full
    return out


def sharded_comp_fcn(state, group=None, hist_fname=None):
    """F(x) of a batched state (coloured probes, Armijo candidates, perturbed iterates) with its
    members sharded over the ranks: rank g evaluates the member block member_block_range(B, g, G) — a
    model year each, no collective on the data path — and the result blocks are all-gathered
    (gather_member_blocks) so that every rank returns the full batched F.  Only the rank's own block of
    the state is touched (a view of `state`; nothing is replicated or re-packed).  Member results do not
    depend on which other members share a launch (tested bit for bit), so the sharded result equals the
    single-GPU one.  Outside a process group this is state.comp_fcn."""
    if not is_sharded(group):
        return state.comp_fcn(None, None, hist_fname)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B = state.members
    width = member_block_width(B, world)
    lo, hi = member_block_range(B, rank, world)
    res = state._like(clone_vals=False)
    local_fcn = None
    if hi > lo:
        local_fcn = state.member_slice(lo, hi).comp_fcn(None, None, hist_fname if lo == 0 else None)
    for ind, tms in enumerate(state.tracer_modules):
        block = tms.vals.new_zeros(tuple(tms.vals.shape[:-1]) + (width,))
        if local_fcn is not None:
            lv = local_fcn.tracer_modules[ind].vals
            block[..., : hi - lo] = lv[..., : hi - lo]
        res.tracer_modules[ind].vals = gather_member_blocks(block, B, group)
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
