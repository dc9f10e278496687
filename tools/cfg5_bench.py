#!/usr/bin/env python
"""BASELINE.json configs[4]: CKKS N = 32768, 512x512 plaintext-diagonal matvec over a batch of encrypted
vectors, sharded BY DIAGONALS across the ranks: every rank holds all ciphertexts, owns a
contiguous range of giant steps, and the partial ciphertexts are summed with one NCCL uint64 all-reduce,
reduced mod q (hegpu_reduce_fixup) and rescaled.

  torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/cfg5_bench.py [--batch 64] [--steps 5]

Prints one JSON line on rank 0: matvecs/s (max over ranks, CUDA events), the all-reduce share, and the
decrypted error of the first ciphertext against numpy."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import hegpu_loader

    hg = hegpu_loader.load()
    from hegpu_b200.client import Client, coeff_modulus_create
    from hegpu_b200.multigpu import allreduce_sum, batch_slice, diag_group, giant_step_range, grid_2d, reduce_scatter_sum

    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--reduce-scatter", action="store_true",
                    help="sum the partials with a reduce-scatter over the batch: every rank rescales (and keeps) its own share")
    ap.add_argument("--diag-ranks", type=int, default=0,
                    help="ranks per diagonal-sharding group (default: all = pure diagonal sharding); the groups split the batch")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, dim, n1, n2, L, scale, B = 32768, 512, 32, 16, 3, 2.0**40, a.batch
    moduli = coeff_modulus_create(N, (60, 40, 40, 60))
    ctx = hg.Context(N, moduli, device=local)
    client = Client(ctx, seed=99)  # same keys and inputs on every rank
    D_ranks = a.diag_ranks or world
    bg, dr, nbg = grid_2d(world, rank, D_ranks)
    group = diag_group(world, rank, D_ranks) if world > 1 else None
    g0, cnt = giant_step_range(n2, D_ranks, dr)
    bfirst, B = batch_slice(a.batch, nbg, bg)  # this batch group's ciphertexts
    steps = list(range(1, n1)) + [g * n1 for g in range(max(g0, 1), g0 + cnt)]
    ctx.load_galois_keys(client.galois_keys_for_steps(steps))
    rng = np.random.default_rng(5)
    M = rng.uniform(-1, 1, (dim, dim))
    V = rng.uniform(-1, 1, (a.batch, dim))[bfirst:bfirst + B]
    slots = N // 2
    r = np.arange(dim)
    rows = np.empty((cnt * n1, slots))
    for g in range(g0, g0 + cnt):
        for k in range(n1):
            d = g * n1 + k
            rows[(g - g0) * n1 + k] = np.roll(np.tile(M[r, (r + d) % dim], slots // dim), g * n1)
    D = ctx.upload_pt_ext(client.encode_many(rows, scale, L, special=True), scale)
    cts = client.encrypt_many(client.encode_many(np.tile(V, (1, slots // dim)), scale, L))
    X = ctx.upload_ct(cts, scale, size_cap=2, L_cap=L)
    part, out = ctx.ct(B, 2, L), ctx.ct(B, 2, L - 1)
    rs = a.reduce_scatter and D_ranks > 1
    if rs:
        mine, out = ctx.ct(B // D_ranks, 2, L), ctx.ct(B // D_ranks, 2, L - 1)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step(timed=False):
        if timed:
            ev[0].record(stream)
        ctx.matvec_bsgs(part, X, D, n1, cnt, rescale=False, dh=True, g_first=g0)
        if timed:
            ev[1].record(stream)
        if rs:
            reduce_scatter_sum(part, mine, D_ranks, group)
        elif D_ranks > 1:
            allreduce_sum(part, D_ranks, group)
        if timed:
            ev[2].record(stream)
        ctx.rescale_to_next(out, mine if rs else part)
        if timed:
            ev[3].record(stream)

    for _ in range(a.warmup):
        step()
    ctx.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tot = ar = 0.0
    for _ in range(a.steps):
        step(timed=True)
        ctx.sync()
        torch.cuda.synchronize()
        tot += ev[0].elapsed_time(ev[3])
        ar += ev[1].elapsed_time(ev[2])
    t = torch.tensor([tot, ar], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot, ar = float(t[0]), float(t[1])
    if rank == 0:
        got = out.download()
        dec = client.decode(client.decrypt(got[0]), out.scale).real[:dim]
        err = float(np.max(np.abs(dec - M @ V[0])))
        print(json.dumps({"config": "cfg5: N=32768 {60,40,40,60}, 512x512, double-hoisted 32x16 sharded by giant steps", "n_gpus": world,
                          "sharding": f"{nbg} batch group(s) x {D_ranks} diagonal rank(s)", "collective": "reduce_scatter" if rs else "all_reduce",
                          "batch": a.batch, "steps": a.steps, "ms_per_step": tot / a.steps, "matvecs_per_s": a.batch * a.steps / (tot * 1e-3),
                          "allreduce_ms_per_step": ar / a.steps, "allreduce_bytes": int(B * 2 * L * N * 8),
                          "max_abs_err_vs_numpy": err, "tolerance": dim * 3.2 * N**1.5 / (8 * scale)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
