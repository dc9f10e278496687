"""Multi-GPU path.

CPU (gloo, world_size 2): the diagonal-sharded algebra -- each rank computes the partial
ciphertext of its giant-step range (oracle), the partials are summed as raw int64 with one
all_reduce and reduced mod q: the result must be bit-identical to the single-process matvec.
GPU: the same on one GPU (ranks emulated sequentially, torch int64 add instead of NCCL) through
hegpu_matvec_bsgs_range + hegpu_reduce_fixup, and with real NCCL when two GPUs are visible."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import hegpu_loader
from fixtures import setup
from oracle import oracle as orc

N, DIM, N1, N2, L, SCALE, B = 4096, 8, 2, 4, 2, 2.0**15, 2
BITS = (36, 36, 37)


def _problem():
    S = setup(N, BITS)
    rng = np.random.default_rng(17)
    from fixtures import rand_residues

    cts = rand_residues(rng, S.moduli[:L], (B, 2), N)
    pts = rand_residues(rng, S.moduli[:L], (N1 * N2,), N)
    steps = list(range(1, N1)) + [g * N1 for g in range(1, N2)]
    gk = S.gk(steps)
    bk = [None] + [gk[orc.galois_elt_from_step(N, k)] for k in range(1, N1)]
    gkeys = [None] + [gk[orc.galois_elt_from_step(N, g * N1)] for g in range(1, N2)]
    return S, cts, pts, gk, bk, gkeys


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hg = hegpu_loader.load()
    from hegpu_b200.multigpu import giant_step_range

    S, cts, pts, gk, bk, gkeys = _problem()
    g0, cnt = giant_step_range(N2, world, rank)
    part = S.o.matvec_bsgs(cts, N1, cnt, pts[g0 * N1:(g0 + cnt) * N1], bk, gkeys[g0:g0 + cnt], hoist=True, lazy=False,
                           rescale=False, g_first=g0)
    t = torch.from_numpy(part.view(np.int64))
    # reduce-scatter over the batch dimension (multigpu.reduce_scatter_sum): rank r receives its share of the sum
    share = torch.empty(t.numel() // world, dtype=torch.int64)
    dist.reduce_scatter_tensor(share, t.reshape(-1).clone(), op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)  # uint64 sum == int64 sum bit for bit
    per = B // world
    assert torch.equal(share, t[rank * per:(rank + 1) * per].reshape(-1))
    summed = t.numpy().view(np.uint64)
    for i, qi in enumerate(S.moduli[:L]):
        summed[:, :, i, :] %= np.uint64(qi)  # what hegpu_reduce_fixup does on the device
    res = np.stack([S.o.rescale(summed[b]) for b in range(B)])
    if rank == 0:
        q.put(res)
    dist.destroy_process_group()


def test_diag_sharded_sum_equals_single_process_gloo():
    hg = hegpu_loader.load()
    from hegpu_b200.multigpu import giant_step_range

    assert [giant_step_range(8, 3, r) for r in range(3)] == [(0, 3), (3, 3), (6, 2)]
    assert [giant_step_range(2, 4, r) for r in range(4)] == [(0, 1), (1, 1), (2, 0), (2, 0)]
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    S, cts, pts, gk, bk, gkeys = _problem()
    want = S.o.matvec_bsgs(cts, N1, N2, pts, bk, gkeys, hoist=True, lazy=False)
    assert np.array_equal(got, want)


def _gloo_worker_2d(rank, world, diag_ranks, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hegpu_loader.load()
    from hegpu_b200.multigpu import batch_slice, diag_group, giant_step_range, grid_2d

    S, cts, pts, gk, bk, gkeys = _problem()
    bg, dr, nbg = grid_2d(world, rank, diag_ranks)
    group = diag_group(world, rank, diag_ranks)
    g0, cnt = giant_step_range(N2, diag_ranks, dr)
    b0, bn = batch_slice(B, nbg, bg)
    part = S.o.matvec_bsgs(cts[b0:b0 + bn], N1, cnt, pts[g0 * N1:(g0 + cnt) * N1], bk, gkeys[g0:g0 + cnt], hoist=True, lazy=False,
                           rescale=False, g_first=g0)
    t = torch.from_numpy(part.view(np.int64))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)  # over the batch group only
    summed = t.numpy().view(np.uint64)
    for i, qi in enumerate(S.moduli[:L]):
        summed[:, :, i, :] %= np.uint64(qi)
    res = np.stack([S.o.rescale(summed[b]) for b in range(bn)])
    if dr == 0:
        q.put((b0, res))
    dist.destroy_process_group()


def test_hybrid_batch_by_diagonal_grid_gloo():
    """2 batch groups x 2 diagonal ranks (world 4): every batch group sums its own partials; together the
    groups produce the single-process result bit for bit."""
    hegpu_loader.load()
    from hegpu_b200.multigpu import batch_slice, grid_2d

    assert [grid_2d(8, r, 2) for r in (0, 1, 2, 7)] == [(0, 0, 4), (0, 1, 4), (1, 0, 4), (3, 1, 4)]
    assert [batch_slice(64, 4, g) for g in range(4)] == [(0, 16), (16, 16), (32, 16), (48, 16)]
    with pytest.raises(ValueError):
        grid_2d(8, 0, 3)
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_gloo_worker_2d, args=(r, 4, 2, port, q)) for r in range(4)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    S, cts, pts, gk, bk, gkeys = _problem()
    want = S.o.matvec_bsgs(cts, N1, N2, pts, bk, gkeys, hoist=True, lazy=False)
    assert np.array_equal(np.concatenate([got[0], got[1]]), want)


@pytest.mark.gpu
def test_diag_sharded_emulated_on_one_gpu():
    hg = hegpu_loader.load()
    from hegpu_b200.multigpu import as_torch_i64, giant_step_range

    S, cts, pts, gk, bk, gkeys = _problem()
    ctx = hg.Context(N, S.moduli)
    ctx.load_galois_keys(gk)
    X = ctx.upload_ct(cts, SCALE, size_cap=2, L_cap=L)
    world = 3
    parts = []
    for r in range(world):
        g0, cnt = giant_step_range(N2, world, r)
        D = ctx.upload_pt(pts[g0 * N1:(g0 + cnt) * N1], SCALE, L_cap=L)
        p = ctx.ct(B, 2, L)
        ctx.matvec_bsgs(p, X, D, N1, cnt, rescale=False, hoist=True, lazy=False, g_first=g0)
        want_part = S.o.matvec_bsgs(cts, N1, cnt, pts[g0 * N1:(g0 + cnt) * N1], bk, gkeys[g0:g0 + cnt], hoist=True, lazy=False,
                                    rescale=False, g_first=g0)
        assert np.array_equal(p.download(), want_part)
        parts.append(p)
    ctx.sync()
    acc = as_torch_i64(parts[0])
    for p in parts[1:]:
        acc += as_torch_i64(p)  # stands in for the NCCL uint64 sum
    torch.cuda.synchronize()
    out, out2 = ctx.ct(B, 2, L - 1), ctx.ct(B, 2, L - 1)
    ctx.rescale_sum_to_next(out2, parts[0], world)  # fix-up folded into the rescale's loads (raw sums in)
    ctx.reduce_fixup(parts[0], world)
    ctx.rescale_to_next(out, parts[0])
    want = S.o.matvec_bsgs(cts, N1, N2, pts, bk, gkeys, hoist=True, lazy=False)
    assert np.array_equal(out.download(), want)
    assert np.array_equal(out2.download(), want)


@pytest.mark.gpu
def test_diag_sharded_double_hoisted_emulated_on_one_gpu():
    """Same with HEGPU_MATVEC_DH partials: every rank's partial is bit-identical to the oracle's
    sharded restatement, and so is the reduced + rescaled sum."""
    hg = hegpu_loader.load()
    from fixtures import rand_residues
    from hegpu_b200.multigpu import as_torch_i64, giant_step_range

    S, cts, pts, gk, bk, gkeys = _problem()
    rng = np.random.default_rng(18)
    ptsx = np.concatenate([pts, rand_residues(rng, [S.moduli[-1]], (N1 * N2,), N)], axis=1)  # + limb mod P
    ctx = hg.Context(N, S.moduli)
    ctx.load_galois_keys(gk)
    X = ctx.upload_ct(cts, SCALE, size_cap=2, L_cap=L)
    world = 2
    parts, want_sum = [], None
    for r in range(world):
        g0, cnt = giant_step_range(N2, world, r)
        sl = np.ascontiguousarray(ptsx[g0 * N1:(g0 + cnt) * N1])
        D = ctx.upload_pt_ext(sl, SCALE)
        p = ctx.ct(B, 2, L)
        ctx.matvec_bsgs(p, X, D, N1, cnt, rescale=False, dh=True, g_first=g0)
        want_part = S.o.matvec_bsgs(cts, N1, cnt, sl, bk, gkeys[g0:g0 + cnt], dh=True, rescale=False, g_first=g0)
        assert np.array_equal(p.download(), want_part)
        want_sum = want_part if want_sum is None else np.stack([S.o.add(want_sum[b], want_part[b]) for b in range(B)])
        parts.append(p)
    ctx.sync()
    acc = as_torch_i64(parts[0])
    acc += as_torch_i64(parts[1])
    torch.cuda.synchronize()
    ctx.reduce_fixup(parts[0], world)
    out = ctx.ct(B, 2, L - 1)
    ctx.rescale_to_next(out, parts[0])
    assert np.array_equal(out.download(), np.stack([S.o.rescale(want_sum[b]) for b in range(B)]))


def _nccl_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    hg = hegpu_loader.load()
    from hegpu_b200.multigpu import giant_step_range, matvec_bsgs_diag_sharded

    S, cts, pts, gk, bk, gkeys = _problem()
    ctx = hg.Context(N, S.moduli, device=rank)
    ctx.load_galois_keys(gk)
    g0, cnt = giant_step_range(N2, world, rank)
    X = ctx.upload_ct(cts, SCALE, size_cap=2, L_cap=L)
    D = ctx.upload_pt(pts[g0 * N1:(g0 + cnt) * N1], SCALE, L_cap=L)
    out = ctx.ct(B, 2, L - 1)
    matvec_bsgs_diag_sharded(ctx, out, X, D, N1, N2, rank, world, hoist=True)
    # the same sum as a reduce-scatter over the batch: this rank ends with ciphertexts [rank*B/world, ...) only
    from hegpu_b200.multigpu import reduce_scatter_sum

    part, mine, out_rs = ctx.ct(B, 2, L), ctx.ct(B // world, 2, L), ctx.ct(B // world, 2, L - 1)
    ctx.matvec_bsgs(part, X, D, N1, cnt, rescale=False, hoist=True, lazy=False, g_first=g0)
    reduce_scatter_sum(part, mine, world, fixup=False)
    ctx.rescale_sum_to_next(out_rs, mine, world)
    per = B // world
    full = out.download()
    assert np.array_equal(out_rs.download(), full[rank * per:(rank + 1) * per])
    if rank == 0:
        q.put(full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_diag_sharded_nccl_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    S, cts, pts, gk, bk, gkeys = _problem()
    want = S.o.matvec_bsgs(cts, N1, N2, pts, bk, gkeys, hoist=True, lazy=False)
    assert np.array_equal(got, want)
