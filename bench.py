#!/usr/bin/env python
"""bench.py -- headline benchmark: CKKS matvecs/s, N = 2^14, 128x128 plaintext matrix x
encrypted vector, double-hoisted BSGS 32x4, batch of 256 ciphertexts per step (BASELINE.json
configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                  [--mode dh|hoist|exact] [--n1 N1] [--batch B] [--no-cfg5] [--no-micro] [--diag-ranks D]

One process per GPU (torchrun for N > 1).  A step is one pass of hegpu_matvec_bsgs over a
batch of 256 encrypted vectors on every rank (batch sharding, no data-path collective ->
weak scaling).  Prints ONE JSON line (rank 0).

  value     matvecs/s with inputs resident in HBM (CUDA events on the context's stream)
  e2e       same through the C ABI with pinned HOST buffers: upload of the input ciphertexts +
            matvec + download of the results inside the timed region
  roofline  the HBM-bound kernel family the north star quotes (NTT / rescale transforms: per-launch CUDA
            events in a second pass of the same steps, algorithmic bytes per DESIGN.md) against
            MEASURED_PEAKS.json, plus `dominant_kernel`: the fused multiply-accumulate kernel against the
            INTEGER / FP64 issue roofline measured in the same run (hegpu_pipe_peak)
  ntt_micro / rescale_micro   stand-alone transforms of 4096 limbs / rescale of 512 ciphertexts (N = 1 only)
  diag_sharded   BASELINE.json configs[4] (N = 32768, 512x512, 1024 ciphertexts in total), giant steps sharded
            over the ranks of a diagonal group, partial ciphertexts summed with an NCCL uint64 reduce-scatter
  copy_micro     pure pinned H2D + D2H copies of the e2e step's bytes on all ranks at once (the host-side ceiling)
  cpu_baseline   the CPU oracle (a port of SEAL 4.1's algorithms; SEAL itself is not installable here)
            running the same BSGS matvec on all host cores, and on one thread, N = 1 only

--impl reference times that CPU port alone on the same config (the reference's evaluator
is Microsoft SEAL, an absent external dependency, so `kind` is "port").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(N=16384, bits=(60, 40, 40, 60), dim=128, n1=32, n2=4, batch=256, scale=2.0**40, L=3, mode="dh")
CFG5 = dict(N=32768, bits=(60, 40, 40, 60), dim=512, n1=32, n2=16, total_batch=1024, scale=2.0**40, L=3)
MODES = {
    "exact": "no flags: the exact chain of SEAL primitives (rotate_vector, multiply_plain, add, rescale_to_next per diagonal)",
    "hoist": "HEGPU_MATVEC_HOIST|LAZY (hoisted baby steps, one mod-down for the giant steps)",
    "dh": "HEGPU_MATVEC_DH (double-hoisted: baby rotations stay in the extended basis, one mod-down per giant step)",
}
METRIC = "CKKS matvecs/s (N=2^14, 128x128)"


# ------------------------------------------------------------------ helpers
def tolerance(mode, n_terms, N, scale):
    """Decryption tolerance stated in DESIGN.md ("Tolerance"), <= ~10x the measured error: the exact chain of SEAL
    primitives carries the random-walk key-switch noise of non-centred RNS digits (n * sigma * N^1.5 / (8 * scale));
    the hoisted modes multiply every key-switched term by a diagonal first (2400 * sqrt(n * N) / scale)."""
    if mode == "exact":
        return n_terms * 3.2 * N**1.5 / (8.0 * scale)
    return 2400.0 * (n_terms * N) ** 0.5 / scale


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed regions (B200_PROFILING.md): one
    `nvidia-smi -lms 50` child that streams CSV rows while the steps run."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc = device, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def summary(self):
        rows = []
        if self.proc is not None:
            try:
                self.proc.terminate()
                out, _ = self.proc.communicate(timeout=5)
                rows = [[x.strip() for x in ln.split(",")] for ln in out.splitlines() if ln.count(",") >= 6]
            except Exception:
                rows = []
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        power = [float(r[2]) if r[2].replace(".", "").isdigit() else 0.0 for r in rows]
        pmax = max(power) if power else 0.0
        loaded = [r for r, p in zip(rows, power) if p >= 0.6 * pmax] or rows  # samples taken while the steps ran
        sm = sorted(float(r[0]) for r in loaded if r[0].replace(".", "").isdigit())
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(r[col].lower() == "active" for r in rows if len(r) > col):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "samples": len(rows), "samples_under_load": len(loaded), "power_w_max": pmax or None,
                "window": "resident + end-to-end timed regions, 50 ms period; sm_mhz = median of the samples under load"}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def diag_rows(M, n1, g0, cnt, slots):
    """pre-rotated plaintext diagonals of giant steps g0 .. g0+cnt-1 (BSGS: diagonal g*n1+k rotated right by g*n1)"""
    dim = M.shape[0]
    rows = np.empty((cnt * n1, slots))
    r = np.arange(dim)
    for g in range(g0, g0 + cnt):
        for b in range(n1):
            d = g * n1 + b
            rows[(g - g0) * n1 + b] = np.roll(np.tile(M[r, (r + d) % dim], slots // dim), g * n1)
    return rows


def synth_problem(client, seed):
    """128x128 matrix -> 128 pre-rotated plaintext diagonals; `batch` encrypted vectors."""
    c = CFG
    rng = np.random.default_rng(seed)
    dim, slots = c["dim"], c["N"] // 2
    M = rng.uniform(-1, 1, (dim, dim))
    V = rng.uniform(-1, 1, (c["batch"], dim))
    pts = client.encode_many(diag_rows(M, c["n1"], 0, c["n2"], slots), c["scale"], c["L"], special=c["mode"] == "dh")
    plains = client.encode_many(np.tile(V, (1, slots // dim)), c["scale"], c["L"])
    cts = client.encrypt_many(plains)
    return M, V, pts, cts


def workload_name():
    c = CFG
    return (f"cfg2: CKKS N={c['N']} {{60,40,40,60}}, {c['dim']}x{c['dim']} plaintext diagonals x encrypted vector, "
            f"BSGS {c['n1']}x{c['n2']}")


def config_record(world):
    """Same keys and values in both arms (the reference arm times a bounded sample of this workload)."""
    c = CFG
    return {"workload": workload_name(), "batch_per_gpu_per_step": c["batch"],
            "parallelism": f"batch-sharded x{world}, keys and diagonals replicated",
            "mode": MODES[c["mode"]] + "; the CPU arm runs the same algorithm",
            "l2": "no flush: each step streams > 2 GB of ciphertexts, lifted digits, keys, diagonals and scratch per GPU, far beyond the 126 MB L2"}


def host_threads():
    """All the host cores this process may use.  (torchrun exports OMP_NUM_THREADS=1, so
    omp_get_max_threads() would say 1; the oracle takes the thread count explicitly.)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def bind_to_gpu_numa_node(device):
    """Pin this rank's host threads to the CPU cores next to its GPU (NVML's ideal affinity), so that the
    pinned staging buffers it allocates afterwards are first-touched on that NUMA node: with 8 ranks the
    host<->device copies of the end-to-end path otherwise cross the socket interconnect."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def rotation_steps():
    return list(range(1, CFG["n1"])) + [g * CFG["n1"] for g in range(1, CFG["n2"])]


def oracle_kwargs():
    m = CFG["mode"]
    return dict(fast=m == "hoist", dh=m == "dh")


# ------------------------------------------------------------------ reference arm (CPU)
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc

    c = CFG
    threads = host_threads()
    moduli = orc.coeff_modulus_create(c["N"], c["bits"])
    o = orc.Oracle(c["N"], moduli)
    rng = np.random.default_rng(1)
    s = o.sample_secret(7)
    gk = o.gen_galois_keys_for_steps(11, s, rotation_steps())
    bk = [None] + [gk[orc.galois_elt_from_step(c["N"], k)] for k in range(1, c["n1"])]
    gkeys = [None] + [gk[orc.galois_elt_from_step(c["N"], g * c["n1"])] for g in range(1, c["n2"])]
    mods = moduli[: c["L"]]
    dh = c["mode"] == "dh"
    pmods = mods + [moduli[-1]] if dh else mods
    pts = np.empty((c["dim"], len(pmods), c["N"]), dtype=np.uint64)
    for i, q in enumerate(pmods):
        pts[:, i, :] = rng.integers(0, q, size=(c["dim"], c["N"]), dtype=np.uint64)
    per_step = threads  # one ciphertext per host thread per step (bounded sample of the batch)
    cts = np.empty((per_step, 2, c["L"], c["N"]), dtype=np.uint64)
    for i, q in enumerate(mods):
        cts[:, :, i, :] = rng.integers(0, q, size=(per_step, 2, c["N"]), dtype=np.uint64)
    for _ in range(args.warmup):
        o.matvec_bsgs(cts, c["n1"], c["n2"], pts, bk, gkeys, threads=threads, **oracle_kwargs())
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.matvec_bsgs(cts, c["n1"], c["n2"], pts, bk, gkeys, threads=threads, **oracle_kwargs())
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = (f"{per_step} ciphertexts per step (1 per host thread) of the {c['batch']}-ciphertext batch, same BSGS {c['n1']}x{c['n2']} matvec "
              f"({c['mode']}); CPU port of SEAL 4.1's algorithms (oracle/), real SEAL is not installable here")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "matvecs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": config_record(args.gpus),
        "cpu_baseline": {"value": value, "unit": "matvecs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "matvecs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ GPU arm: sub-benchmarks
def events_ms(torch, stream, ctx, fn, iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.sync()
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    ctx.sync()
    return e0.elapsed_time(e1) / iters


def ntt_and_rescale_micro(torch, hg, ctx, stream, moduli, peak):
    """BASELINE.md row mu: 4096 limb transforms of the {60,40,40,60} chain at N = 16384 (16*N algorithmic bytes per
    limb), and rescale_to_next of 512 level-3 ciphertexts (per-kernel events: half_intt + rescale_ntt)."""
    c = CFG
    n, K, count = c["N"], len(moduli), 4096
    rng = np.random.default_rng(0)
    host = np.empty((count, n), dtype=np.uint64)
    for i in range(K):
        host[i::K] = rng.integers(0, moduli[i], size=host[i::K].shape, dtype=np.uint64)
    d = torch.from_numpy(host.view(np.int64)).cuda()
    out = {"n": n, "limbs": count, "chain": list(c["bits"]), "bytes_per_limb": 16 * n}
    for name, fn in (("fwd", ctx.ntt_forward_device), ("inv", ctx.ntt_inverse_device)):
        call = lambda: fn(d.data_ptr(), count, 0, K)  # noqa: E731
        for _ in range(3):
            call()
        ms = events_ms(torch, stream, ctx, call, 20)
        gbs = count * 16 * n / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "GBps": gbs, "frac_of_hbm_peak": gbs / peak}
    out["roundtrip_ok"] = bool(np.array_equal(d.cpu().numpy().view(np.uint64), host))  # 23 fwd + 23 inv
    del d
    B, L = 512, c["L"]
    cts = np.empty((B, 2, L, n), dtype=np.uint64)
    for i in range(L):
        cts[:, :, i, :] = rng.integers(0, moduli[i], size=(B, 2, n), dtype=np.uint64)
    X, OUT = ctx.upload_ct(cts, c["scale"] ** 2, size_cap=2, L_cap=L), ctx.ct(B, 2, L - 1)
    call = lambda: ctx.rescale_to_next(OUT, X)  # noqa: E731
    for _ in range(3):
        call()
    ms = events_ms(torch, stream, ctx, call, 20)
    ctx.profile_reset()
    ctx.profile(True)
    for _ in range(5):
        call()
    prof = ctx.profile_read()
    ctx.profile(False)
    by = sum(v["algo_bytes"] for v in prof.values()) / 5
    kms = sum(v["ms"] for v in prof.values()) / 5
    rs = {"ciphertexts": B, "level": L, "ms": ms, "algo_bytes": by, "GBps": by / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": by / (ms * 1e-3) / 1e9 / peak,
          "kernel_ms": kms, "kernels": {k: v["ms"] / 5 for k, v in prof.items() if v["launches"]},
          "bytes": "per ciphertext: INTT of the dropped limb 2*16N + (L-1) fused transforms 2*(L-1)*24N (t, operand, result)"}
    return out, rs


def copy_micro(torch, dist, local, h2d_bytes, d2h_bytes, reps=10):
    """Pure pinned host<->device copies of the e2e step's bytes, both directions at once, all ranks at once: the
    host-side ceiling of the end-to-end number (no kernel runs)."""
    hin = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    hout = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    din = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
    dout = torch.empty(d2h_bytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(k):
        for _ in range(k):
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
        s1.synchronize()
        s2.synchronize()

    run(2)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(reps)
    ms = (time.perf_counter() - t0) * 1e3 / reps
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    world = dist.get_world_size() if dist is not None else 1
    return {"ms_per_step_copies_only": ms, "GBps_per_rank": (h2d_bytes + d2h_bytes) / (ms * 1e-3) / 1e9,
            "GBps_all_ranks": world * (h2d_bytes + d2h_bytes) / (ms * 1e-3) / 1e9,
            "note": "H2D and D2H of one e2e step's bytes on two streams, every rank at the same time, max over ranks"}


def run_cfg5(args, torch, dist, hg, world, rank, local):
    """BASELINE.json configs[4]: N = 32768, 512x512 plaintext diagonals, 1024 encrypted vectors in total,
    double-hoisted BSGS 32x16.  The ranks form world/D batch groups of D diagonal ranks (multigpu.grid_2d): a group
    owns 1024*D/world ciphertexts, its ranks split the 16 giant steps, the partial ciphertexts (level L, before the
    rescale) are summed over the group with ONE NCCL uint64 reduce-scatter over the batch dimension, and every rank
    rescales its own share with the mod-q fix-up of the sum folded into the rescale's loads
    (hegpu_rescale_sum_to_next).  D = 1: no collective (the N = 1 base line)."""
    from hegpu_b200.client import Client, coeff_modulus_create
    from hegpu_b200.multigpu import batch_slice, diag_group, giant_step_range, grid_2d, reduce_scatter_sum

    c = CFG5
    N, dim, n1, n2, L, scale = c["N"], c["dim"], c["n1"], c["n2"], c["L"], c["scale"]
    D = args.diag_ranks or (2 if world >= 2 else 1)
    if world % D:
        D = 1
    moduli = coeff_modulus_create(N, c["bits"])
    ctx = hg.Context(N, moduli, device=local)
    client = Client(ctx, seed=99)  # same keys on every rank
    bg, dr, nbg = grid_2d(world, rank, D)
    group = diag_group(world, rank, D) if world > 1 else None
    g0, cnt = giant_step_range(n2, D, dr)
    bfirst, B = batch_slice(c["total_batch"], nbg, bg)
    steps = list(range(1, n1)) + [g * n1 for g in range(max(g0, 1), g0 + cnt)]
    ctx.load_galois_keys(client.galois_keys_for_steps(steps))
    M = np.random.default_rng(5).uniform(-1, 1, (dim, dim))
    V = np.random.default_rng(1000 + bg).uniform(-1, 1, (B, dim))  # this batch group's vectors (same on its D ranks)
    slots = N // 2
    Dg = ctx.upload_pt_ext(client.encode_many(diag_rows(M, n1, g0, cnt, slots), scale, L, special=True), scale)
    crng = client.rng
    client.rng = np.random.default_rng(2000 + bg)  # the D ranks of a group must hold the SAME ciphertexts
    X = ctx.ct(B, 2, L)
    X.set_meta(2, L, scale)
    for b0 in range(0, B, 128):  # encrypt in chunks: bounded host memory
        b1 = min(B, b0 + 128)
        chunk = client.encrypt_many(client.encode_many(np.tile(V[b0:b1], (1, slots // dim)), scale, L))
        for i in range(b1 - b0):
            X.upload_one(b0 + i, chunk[i])
    client.rng = crng
    part = ctx.ct(B, 2, L)
    share = B // D
    mine = ctx.ct(share, 2, L) if D > 1 else None
    out = ctx.ct(share, 2, L - 1)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step(timed=False):
        if timed:
            ev[0].record(stream)
        ctx.matvec_bsgs(part, X, Dg, n1, cnt, rescale=False, dh=True, g_first=g0)
        if timed:
            ev[1].record(stream)
        if D > 1:
            reduce_scatter_sum(part, mine, D, group, fixup=False)
        if timed:
            ev[2].record(stream)
        if D > 1:
            ctx.rescale_sum_to_next(out, mine, D)
        else:
            ctx.rescale_to_next(out, part)
        if timed:
            ev[3].record(stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        ctx.sync()
        torch.cuda.synchronize()

    step()
    barrier()
    l0 = ctx.launches
    k = max(1, args.cfg5_steps)
    tot = mv = coll = resc = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(k):
        step(timed=True)
    e1.record(stream)
    ctx.sync()
    torch.cuda.synchronize()
    tot = e0.elapsed_time(e1)
    # per-phase split from a second, per-step synchronised pass (the first pass is the timed one)
    for _ in range(k):
        step(timed=True)
        ctx.sync()
        torch.cuda.synchronize()
        mv += ev[0].elapsed_time(ev[1])
        coll += ev[1].elapsed_time(ev[2])
        resc += ev[2].elapsed_time(ev[3])
    launches = (ctx.launches - l0) // (2 * k)
    t = torch.tensor([tot, mv, coll, resc], device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot, mv, coll, resc = (float(x) for x in t)
    rec = None
    if rank == 0:
        got = out.download()
        dec = client.decrypt_decode_many(got[:4], out.scale).real[:, :dim]
        want = V[dr * share: dr * share + 4] @ M.T
        err = float(np.max(np.abs(dec - want)))
        tol = tolerance("dh", dim, N, scale)
        if not err < tol:
            raise SystemExit(f"cfg5: decrypted diagonal-sharded matvec is wrong: max |err| {err:.3e} >= tolerance {tol:.3e}")
        rec = {"workload": "cfg5: CKKS N=32768 {60,40,40,60}, 512x512 plaintext diagonals x 1024 encrypted vectors in total, double-hoisted BSGS 32x16",
               "n_gpus": world, "sharding": f"{nbg} batch group(s) x {D} diagonal rank(s): each rank runs {cnt} of the {n2} giant steps on {B} ciphertexts",
               "collective": f"NCCL uint64 reduce-scatter over the batch inside each group of {D} (torch.distributed)" if D > 1 else "none (one rank per group)",
               "value": c["total_batch"] * k / (tot * 1e-3), "unit": "matvecs/s", "steps": k, "ms_per_step": tot / k,
               "matvec_ms": mv / k, "collective_ms": coll / k, "rescale_with_fixup_ms": resc / k,
               "collective_bytes_in_per_rank": int(B * 2 * L * N * 8) if D > 1 else 0,
               "gpu_launches_per_step": int(launches), "max_abs_err_vs_numpy": err, "tolerance": tol,
               "timing": "CUDA events on the context stream around k back-to-back steps, max over ranks; phase split from a second pass synchronised per step"}
    barrier()
    del ctx
    return rec


# ------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch

    import hegpu_loader

    hg = hegpu_loader.load()
    from hegpu_b200.client import Client, coeff_modulus_create  # package registered by hegpu_loader

    c = CFG
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None  # N = 1 keeps every host core for the CPU baseline
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    moduli = coeff_modulus_create(c["N"], c["bits"])
    ctx = hg.Context(c["N"], moduli, device=local)
    client = Client(ctx, seed=1234)  # same keys on every rank (replicated), different vectors per rank
    gk = client.galois_keys_for_steps(rotation_steps())
    ctx.load_galois_keys(gk)
    M, V, pts, cts = synth_problem(client, seed=0xC0FFEE + 2 + 1000 * rank)
    B = c["batch"]

    mode = c["mode"]
    dh = mode == "dh"
    D = ctx.upload_pt_ext(pts, c["scale"]) if dh else ctx.upload_pt(pts, c["scale"])
    X = ctx.upload_ct(cts, c["scale"], size_cap=2, L_cap=c["L"])
    OUT = ctx.ct(B, 2, c["L"] - 1)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)

    def matvec(out, x):
        ctx.matvec_bsgs(out, x, D, c["n1"], c["n2"], hoist=mode == "hoist", dh=dh)

    def step():
        matvec(OUT, X)

    def barrier():
        if dist is not None:
            dist.barrier()
        ctx.sync()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        ctx.sync()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    # correctness of the measured path: EVERY ciphertext of the step is decrypted and compared with numpy
    step()
    got = OUT.download()
    out_scale = OUT.scale
    dec = client.decrypt_decode_many(got, out_scale).real[:, : c["dim"]]
    full0 = client.decode(client.decrypt(got[0]), out_scale).real[: c["dim"]]  # full-basis decode of ciphertext 0 agrees with the fast path
    tol = tolerance(mode, c["dim"], c["N"], c["scale"])
    errs = np.max(np.abs(dec - V @ M.T), axis=1)
    max_err = float(errs.max())
    if not max_err < tol or not np.max(np.abs(full0 - dec[0])) < 1e-9:
        raise SystemExit(f"decrypted matvec is wrong: max |err| {max_err:.3e} >= tolerance {tol:.3e}")

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launches
    ms = timed(step, args.steps)
    launches = ctx.launches - l0
    value = world * B * args.steps / (ms * 1e-3)

    # ---- e2e: pinned host buffers, H2D + matvec + D2H every step, through the C ABI's asynchronous
    # copies (double-buffered: the upload of step i+1 and the download of step i-1 overlap the
    # matvec of step i).  Timed with the host clock between full synchronisations.
    h_in = [torch.from_numpy(cts.copy()).pin_memory() for _ in range(2)]
    h_out = [torch.empty((B, 2, c["L"] - 1, c["N"]), dtype=torch.int64).pin_memory() for _ in range(2)]
    XS = [ctx.ct(B, 2, c["L"]) for _ in range(2)]
    OS = [ctx.ct(B, 2, c["L"] - 1) for _ in range(2)]

    def e2e_run(steps):
        XS[0].upload_async(h_in[0].data_ptr(), c["scale"], 2, c["L"])
        for i in range(steps):
            cur, nxt = i & 1, (i + 1) & 1
            if i + 1 < steps:
                XS[nxt].upload_async(h_in[nxt].data_ptr(), c["scale"], 2, c["L"])
            if i >= 2:
                OS[cur].copy_wait()  # host buffer h_out[cur] of step i-2 is complete (a consumer would read it here)
            matvec(OS[cur], XS[cur])
            OS[cur].download_async(h_out[cur].data_ptr())
        for o_ in OS:
            o_.copy_wait()
        ctx.sync()

    e2e_run(3)
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    ms_e2e = (time.perf_counter() - t0) * 1e3
    if dist is not None:
        t = torch.tensor([ms_e2e], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    barrier()
    clocks = sampler.summary()
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    for ho in h_out:
        assert np.array_equal(ho.numpy().view(np.uint64), got), "e2e path result differs from the resident path"
    h2d_bytes, d2h_bytes = int(cts.nbytes), int(h_out[0].numel() * 8)
    cm = copy_micro(torch, dist, local, h2d_bytes, d2h_bytes)
    del h_in, h_out, XS, OS

    # ---- per-kernel pass (same steps, events around every launch) -> rooflines
    ctx.profile_reset()
    ctx.profile(True)
    for _ in range(args.steps):
        step()
    prof = ctx.profile_read()
    ctx.profile(False)
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
    dom = max(prof, key=lambda k: prof[k]["ms"])
    peak, peak_src = measured_peak_gbs()
    kernels = {k: {"ms_per_step": v["ms"] / args.steps, "share": v["ms"] / tot_ms, "launches_per_step": v["launches"] // args.steps,
                   "algo_GBps": (v["algo_bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] else 0.0}
               for k, v in prof.items() if v["launches"]}
    ntt_keys = [k for k in prof if "ntt" in k and prof[k]["launches"]]
    ntt_ms = sum(prof[k]["ms"] for k in ntt_keys)
    ntt_bytes = sum(prof[k]["algo_bytes"] for k in ntt_keys)
    ntt_launches = sum(prof[k]["launches"] for k in ntt_keys)
    ntt_gbs = ntt_bytes / (ntt_ms * 1e-3) / 1e9 if ntt_ms else 0.0
    # DRAM traffic per launch comes from the committed ncu capture of this configuration (profiles/), not from this run
    profiled, capture_config = {}, None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            # the captures are per LAUNCH: one launch covers one execution slot's chunk (128 ciphertexts at batch 128 with
            # one slot and at batch 256 with two), so they hold for any batch with the same mode and BSGS split
            if str(tj.get("config", "")).startswith(f"{mode}-{c['n1']}x{c['n2']}-"):
                profiled = tj.get("kernels", {})
                capture_config = tj.get("config")
        except Exception:
            pass
    # arithmetic ceilings measured in this run: the product's own butterfly / multiply-accumulate sequences on
    # register operands only, every SM busy (hegpu_pipe_peak)
    pk = {name: ctx.pipe_peak(kind) for kind, name in enumerate(("imad_wide", "dfma", "imad", "mac128", "bfly60", "bfly_fp64"))}
    logn = c["N"].bit_length() - 1
    bfly_to_gbs = 32.0 / logn / 1e9  # a limb is N/2 * logN butterflies for 16*N algorithmic bytes
    roofline = {
        "bound": "hbm", "kernel": "ntt family (" + ", ".join(sorted(ntt_keys)) + ")", "achieved": ntt_gbs, "peak": peak, "unit": "GB/s",
        "frac": ntt_gbs / peak, "peak_source": peak_src,
        "traffic": (sum(profiled[k]["dram_bytes_per_launch"] for k in profiled if "ntt" in k) or None) if profiled else None,
        "traffic_source": f"from_profile (profiles/ncu_traffic.json, capture config {capture_config}: ncu --set full captures of the NTT kernels named there, per launch; not measured in this run)" if profiled else None,
        "avg_launch_ms": ntt_ms / max(ntt_launches, 1), "share_of_step": ntt_ms / tot_ms,
        "arith_ceiling_GBps": {"60bit_butterfly": pk["bfly60"] * bfly_to_gbs, "fp64_butterfly": pk["bfly_fp64"] * bfly_to_gbs,
                               "note": "butterflies/s of the product's own butterfly code on register operands (no loads, exchanges, barriers) "
                                       "converted at N/2*logN butterflies per 16*N bytes: what the issue slots allow before HBM matters"},
        "kernels": kernels,
    }
    if dh and "dh_inner" in prof and prof["dh_inner"]["launches"]:
        d = prof["dh_inner"]
        per_unit = (c["n1"] - 1) * 2 * c["L"] + c["n1"] * 2 * c["n2"]  # key products + diagonal products per (ciphertext, ext limb, coefficient)
        lim60 = sum(1 for q in list(moduli[: c["L"]]) + [moduli[-1]] if q >= (1 << 40))
        lim40 = c["L"] + 1 - lim60
        units = B * c["N"] * args.steps
        p60, p40 = units * lim60 * per_unit, units * lim40 * per_unit
        t_min = p60 / pk["mac128"] + p40 * 4 / pk["dfma"]  # FP64 policy: 4 DFMA per product
        secs = d["ms"] * 1e-3
        hbm = d["algo_bytes"] / secs / 1e9
        roofline["dominant_kernel"] = {
            "kernel": "dh_inner", "bound": "int", "unit": "G products/s (64x64->128-bit multiply-accumulates)",
            "achieved": (p60 + p40) / secs / 1e9, "peak": (p60 + p40) / t_min / 1e9, "frac": t_min / secs,
            "products_per_unit": per_unit, "limbs_int64_policy": lim60, "limbs_fp64_policy": lim40,
            "peak_definition": "time floor = products on >= 2^40 limbs / measured mac128 rate + 4 DFMA per product on the limbs below 2^40 / measured DFMA rate "
                               "(the two policies run in different CTAs; if both pipes overlapped perfectly the floor would be the max, not the sum)",
            "avg_launch_ms": d["ms"] / d["launches"], "share_of_step": d["ms"] / tot_ms,
            "hbm_view": {"achieved_GBps": hbm, "frac": hbm / peak, "traffic": profiled.get("dh_inner", {}).get("dram_bytes_per_launch"),
                         "traffic_source": "from_profile" if "dh_inner" in profiled else None,
                         "note": "not HBM-bound by construction: the fusion removed ~10 GB of traffic per step"},
        }
    roofline["pipe_peaks"] = {"unit": "thread-level operations/s, every SM busy, measured in this run", **pk}
    roofline["dominant"] = dom

    line = {
        "metric": METRIC, "value": value, "unit": "matvecs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": config_record(world),
        "check": {"tolerance": tol, "max_abs_err_vs_numpy": max_err, "ciphertexts_decrypted": int(B),
                  "tolerance_rule": "exact: n*3.2*N^1.5/(8*scale); hoisted/dh: 2400*sqrt(n*N)/scale (DESIGN.md, <= ~10x measured)"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "matvecs/s", "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes,
                "pipeline": "double-buffered hegpu_ct_upload_async / download_async, host-clock timed",
                "host_cores_bound_per_rank": numa, "copy_micro": cm},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }

    # ---- opt-in integer-MMA form of the fused inner sums (HEGPU_MATVEC_IMMA; experimental, NOT the headline: the north
    # star keeps this path off the tensor cores): same bits as the default kernel, timed the same way
    if dh and not args.no_imma:
        OUT2 = ctx.ct(B, 2, c["L"] - 1)

        def step_imma():
            ctx.matvec_bsgs(OUT2, X, D, c["n1"], c["n2"], dh=True, imma=True)

        for _ in range(3):
            step_imma()
        same = bool(np.array_equal(OUT2.download(), got))
        ms_i = timed(step_imma, args.steps)
        line["imma_mode"] = {"value": world * B * args.steps / (ms_i * 1e-3), "unit": "matvecs/s", "ms_per_step": ms_i / args.steps,
                             "bit_identical_to_default": same,
                             "what": "HEGPU_MATVEC_IMMA: the fused inner sums as exact 8-bit-limb integer matrix products on the warp-level integer MMA units "
                                     "(mma.sync m16n8k32 u8), diagonals pre-multiplied with the baby-step keys once; opt-in, experimental, not the headline"}
        del OUT2

    # ---- stand-alone transforms and rescale (N = 1 only)
    if world == 1 and not args.no_micro:
        line["ntt_micro"], line["rescale_micro"] = ntt_and_rescale_micro(torch, hg, ctx, stream, moduli, peak)

    # ---- CPU baseline on rank 0, N = 1 only: the oracle on a bounded sample of the same batch
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc

        threads = host_threads()
        o = orc.Oracle(c["N"], moduli)
        bk = [None] + [gk[ctx.galois_elt_from_step(k)] for k in range(1, c["n1"])]
        gkeys = [None] + [gk[ctx.galois_elt_from_step(g * c["n1"])] for g in range(1, c["n2"])]
        sample_n = min(B, max(threads, 1) * 4)
        half = sample_n // 2  # half of the sample from the start and half from the end of the batch (both execution slots)
        pick = np.r_[0:half, B - (sample_n - half):B] if sample_n < B else np.arange(B)
        sub = np.ascontiguousarray(cts[pick])
        passes, total, best = 0, 0.0, None
        while total < 10.0 and passes < 64:  # about 10 s of CPU work on the sample (the host is shared and noisy)
            t0 = time.perf_counter()
            ref_out = o.matvec_bsgs(sub, c["n1"], c["n2"], pts, bk, gkeys, threads=threads, **oracle_kwargs())
            dt = time.perf_counter() - t0
            total += dt
            passes += 1
            best = dt if best is None else min(best, dt)
        line["cpu_baseline"] = {"value": sample_n * passes / total, "unit": "matvecs/s", "cores": threads, "kind": "port",
                                "best_pass_value": sample_n / best,
                                "sample": f"{sample_n} of the {B} ciphertexts of one step (first and last {half}), same keys/diagonals and algorithm, "
                                          f"{passes} passes, {total:.1f} s wall in total",
                                "bit_exact_vs_gpu": bool(np.array_equal(ref_out, got[pick]))}
        # the reference itself is single-threaded (BASELINE.md tier B1): the same port on ONE thread
        one = np.ascontiguousarray(cts[:2])
        passes1, total1 = 0, 0.0
        while total1 < 3.0 and passes1 < 16:
            t0 = time.perf_counter()
            ref1 = o.matvec_bsgs(one, c["n1"], c["n2"], pts, bk, gkeys, threads=1, **oracle_kwargs())
            total1 += time.perf_counter() - t0
            passes1 += 1
        line["cpu_baseline_1thread"] = {"value": 2 * passes1 / total1, "unit": "matvecs/s", "cores": 1, "kind": "port",
                                        "sample": f"ciphertexts 0-1 of the step, {passes1} passes, {total1:.1f} s wall",
                                        "bit_exact_vs_gpu": bool(np.array_equal(ref1, got[:2]))}
    del ctx, X, OUT, D

    # ---- configs[4]: diagonal-sharded 512x512 matvec over 1024 ciphertexts with the NCCL sum
    if not args.no_cfg5:
        rec = run_cfg5(args, torch, dist, hg, world, rank, local)
        if rank == 0:
            line["diag_sharded"] = rec
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="hegpu", choices=["hegpu", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the diagonal-sharded configs[4] sub-record")
    ap.add_argument("--no-micro", action="store_true", help="skip the stand-alone NTT / rescale records")
    ap.add_argument("--no-imma", action="store_true", help="skip the opt-in integer-MMA sub-record")
    ap.add_argument("--cfg5-steps", type=int, default=3)
    ap.add_argument("--diag-ranks", type=int, default=0, help="configs[4]: ranks per diagonal group (default 2 when N >= 2)")
    ap.add_argument("--mode", default=CFG["mode"], choices=sorted(MODES))
    ap.add_argument("--n1", type=int, default=None, help="baby steps (n1 * n2 = 128)")
    ap.add_argument("--batch", type=int, default=None, help="ciphertexts per GPU per step")
    args = ap.parse_args()
    CFG["mode"] = args.mode
    if args.batch:
        CFG["batch"] = args.batch
    if args.mode in ("hoist", "exact") and not args.n1:
        args.n1 = 16  # the single-hoisting composite takes at most 16 baby steps; 16x8 is also the exact chain's optimum
    if args.n1:
        assert CFG["dim"] % args.n1 == 0
        CFG["n1"], CFG["n2"] = args.n1, CFG["dim"] // args.n1
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
