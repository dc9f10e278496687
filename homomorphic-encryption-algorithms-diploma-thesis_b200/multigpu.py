"""Multi-GPU plumbing: one process per GPU, torch.distributed over NCCL/NVLink.

The path shards by independent units.  Three shardings are provided:

* batch sharding -- every rank runs the whole matvec on its own slice of the ciphertext batch;
  keys and diagonals are replicated, there is no data-path collective (bench.py, "weak");
* hybrid -- grid_2d: batch groups of diag_ranks ranks each, diagonal sharding inside a group;
* diagonal sharding -- rank r owns a contiguous range of giant steps (its n1*cnt diagonals and
  the Galois keys of those giant steps), recomputes the shared baby-step rotations locally and
  produces a partial ciphertext at level L *before* rescale.  The partials are summed with ONE
  collective (uint64 sum == int64 sum bit for bit; residues < q <= 2^60 so up to 16 terms
  cannot wrap), reduced back to [0,q) by hegpu_reduce_fixup and rescaled.  Sum-then-reduce
  equals SEAL's chain of add_inplace bit for bit because modular addition is associative on
  canonical residues.
"""
from __future__ import annotations

import numpy as np


def giant_step_range(n2: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous split of the n2 giant steps: (first global giant step, count) of `rank`."""
    base, extra = divmod(n2, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def grid_2d(world: int, rank: int, diag_ranks: int) -> tuple[int, int, int]:
    """Hybrid sharding: the ranks form (world / diag_ranks) batch groups of diag_ranks consecutive
    ranks.  A batch group owns a slice of the ciphertext batch; inside it the giant steps are split
    (giant_step_range(n2, diag_ranks, diag_rank)) and the partials are summed over the group only.
    diag_ranks = world is pure diagonal sharding, diag_ranks = 1 pure batch sharding: the baby-step
    work that diagonal sharding repeats on every rank shrinks with the group size.
    Returns (batch_group, diag_rank, n_batch_groups)."""
    if diag_ranks < 1 or world % diag_ranks:
        raise ValueError("diag_ranks must divide the number of ranks")
    return rank // diag_ranks, rank % diag_ranks, world // diag_ranks


def batch_slice(batch: int, groups: int, group: int) -> tuple[int, int]:
    """Contiguous split of the ciphertext batch over the batch groups: (first, count)."""
    return giant_step_range(batch, groups, group)


def diag_group(world: int, rank: int, diag_ranks: int):
    """The torch.distributed group of this rank's batch group (None when it is the whole world or a
    single rank).  Collective: every rank must call it."""
    import torch.distributed as dist

    if diag_ranks == world or diag_ranks == 1:
        return None
    mine = None
    for g in range(world // diag_ranks):
        grp = dist.new_group(ranks=list(range(g * diag_ranks, (g + 1) * diag_ranks)))
        if rank // diag_ranks == g:
            mine = grp
    return mine


class _CudaArray:
    """__cuda_array_interface__ wrapper so torch can alias library-owned device memory."""

    def __init__(self, ptr: int, nwords: int):
        self.__cuda_array_interface__ = {"shape": (nwords,), "typestr": "<i8", "data": (ptr, False), "version": 3}


def as_torch_i64(ct):
    """Zero-copy int64 view of a CtBatch's whole device buffer [B][size_cap][L_cap][N]."""
    import torch

    ptr, sb, _, _ = ct.device_view()
    return torch.as_tensor(_CudaArray(ptr, ct.batch * sb), device=f"cuda:{ct.ctx.device}")


def allreduce_sum(ct, world: int, group=None, fixup: bool = True):
    """In-place sum of a partial ciphertext batch over all ranks; canonical residues on return, or -- fixup=False --
    the raw uint64 sums (< world * q), to be consumed by Context.rescale_sum_to_next, which reduces while it loads."""
    import torch
    import torch.distributed as dist

    ctx = ct.ctx
    if world > 16:
        raise ValueError("at most 16 partial sums of 60-bit residues fit in 64 bits; reduce hierarchically")
    t = as_torch_i64(ct)
    with torch.cuda.stream(torch.cuda.ExternalStream(ctx.stream, device=ctx.device)):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)  # NCCL over NVLink / NVSwitch
    if fixup:
        ctx.reduce_fixup(ct, world)


def reduce_scatter_sum(part, mine, world: int, group=None, fixup: bool = True):
    """Sum of the partial ciphertext batches over the ranks, scattered over the batch dimension:
    rank r of the group receives ciphertexts [r*B/world, (r+1)*B/world) of the sum in `mine` (canonical residues).
    Half the bytes of an all-reduce on the wire, and every rank rescales only its own share afterwards.
    fixup=False leaves the raw uint64 sums for Context.rescale_sum_to_next."""
    import torch
    import torch.distributed as dist

    ctx = part.ctx
    if world > 16:
        raise ValueError("at most 16 partial sums of 60-bit residues fit in 64 bits; reduce hierarchically")
    if part.batch % world or mine.batch != part.batch // world:
        raise ValueError("the batch must split evenly over the ranks of the group")
    src, dst = as_torch_i64(part), as_torch_i64(mine)
    if dst.numel() * world != src.numel():
        raise ValueError("partial and scattered batches must share their layout (size_cap, L_cap)")
    with torch.cuda.stream(torch.cuda.ExternalStream(ctx.stream, device=ctx.device)):
        dist.reduce_scatter_tensor(dst, src, op=dist.ReduceOp.SUM, group=group)
    _, _, L, scale = part.info()
    mine.set_meta(2, L, scale)
    if fixup:
        ctx.reduce_fixup(mine, world)


def matvec_bsgs_diag_sharded(ctx, out, x, diags_local, n1: int, n2_total: int, rank: int, world: int, hoist: bool = True,
                             partial=None, group=None, dh: bool = False):
    """Diagonal-sharded BSGS matvec.  diags_local holds the n1*cnt pre-rotated diagonals of this
    rank's giant steps (giant_step_range).  Every rank ends with the full result in `out`.
    dh=True: double-hoisted partials (diags_local uploaded with upload_pt_ext); every rank then runs
    its own final mod-down, so the sum equals the unsharded result up to that rounding (and the
    oracle's sharded restatement bit for bit)."""
    g_first, cnt = giant_step_range(n2_total, world, rank)
    _, _, L, _ = x.info()
    part = partial if partial is not None else ctx.ct(x.batch, 2, L)
    if cnt:
        ctx.matvec_bsgs(part, x, diags_local, n1, cnt, rescale=False, hoist=hoist and not dh, lazy=False, g_first=g_first, dh=dh)
    else:  # more ranks than giant steps: contribute zero
        part.upload(np.zeros((x.batch, 2, L, ctx.n), dtype=np.uint64), x.scale * diags_local.scale if diags_local else x.scale)
    if world > 1:
        allreduce_sum(part, world, group, fixup=False)
        ctx.rescale_sum_to_next(out, part, world)  # the mod-q fix-up of the sum rides in the rescale's loads
    else:
        ctx.rescale_to_next(out, part)
    return out
