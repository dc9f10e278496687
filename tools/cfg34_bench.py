#!/usr/bin/env python
"""Timings of BASELINE.json configs[2] and configs[3] (matmuls/s, FFTs/s, key-switches/s) on one
B200, each with the CPU port of the same statement sequence beside it and the per-kernel-family split.

  python tools/cfg34_bench.py [--only bmatmul|matmul|bfft4096|bfft128] [--no-cpu]

  bmatmul    cfg 3 (i):  BatchedMatrix::matmul case B (he_linalg.cpp:943-1006), 64 x 64 encrypted x encrypted, N = 16384
             {60,40,40,60}, SEAL's default power-of-two Galois keys (NAF chains), relinearize + rescale per output
  matmul     cfg 3 (ii): Matrix::matmul (he_linalg.cpp:202-236), 64 x 64, one ciphertext per entry (2 x 4096 inputs)
  bfft4096   cfg 4:      he::fft::bfft stage loop (he_fft.cpp:178-203) over 4096 slots, 12 stages, 13-level chain
             {60, 40 x 12, 60}, 23 rotations per FFT; batch 1 (the reference's shape) and batch 32
  bfft128    the shipped demo's shape (fft.cpp:127-241): n = 128, {60,31,30x9,60}, 7 stages, 13 rotations

Inputs are uniformly random residues (timing only; correctness of every path is in tests/).  One JSON line each."""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def residues(rng, moduli, prefix, n):
    out = np.empty(tuple(prefix) + (len(moduli), n), dtype=np.uint64)
    for i, q in enumerate(moduli):
        out[..., i, :] = rng.integers(0, q, size=tuple(prefix) + (n,), dtype=np.uint64)
    return out


def naf(v):
    out, i = [], 0
    while v:
        if v & 1:
            z = 2 - (v & 3)
            out.append(z << i)
            v -= z
        v >>= 1
        i += 1
    return out


def timed(torch, ctx, stream, fn, iters, warm=1):
    for _ in range(warm):
        fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    ctx.sync()
    return e0.elapsed_time(e1) / iters


def kernel_split(ctx, fn):
    ctx.profile_reset()
    ctx.profile(True)
    fn()
    prof = ctx.profile_read()
    ctx.profile(False)
    tot = sum(v["ms"] for v in prof.values()) or 1.0
    ntt_ms = sum(v["ms"] for k, v in prof.items() if "ntt" in k)
    ntt_b = sum(v["algo_bytes"] for k, v in prof.items() if "ntt" in k)
    return {"kernel_ms": tot, "share": {k: round(v["ms"] / tot, 4) for k, v in prof.items() if v["launches"]},
            "ntt_family_GBps": ntt_b / (ntt_ms * 1e-3) / 1e9 if ntt_ms else None, "launches": int(sum(v["launches"] for v in prof.values()))}


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    import torch

    import hegpu_loader

    hg = hegpu_loader.load()
    from hegpu_b200.client import Client, coeff_modulus_create
    from oracle import oracle as orc

    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    threads = max(1, len(os.sched_getaffinity(0)))
    peak = peak_gbs()

    def emit(rec):
        rec["hbm_peak_GBps"] = peak
        if rec.get("ntt_family_GBps"):
            rec["ntt_family_frac_of_hbm_peak"] = rec["ntt_family_GBps"] / peak
        print(json.dumps(rec), flush=True)

    # ------------------------------------------------------------ cfg 3 (i) and (ii)
    if a.only in ("", "bmatmul", "matmul"):
        n, dim, L, sc = 16384, 64, 3, 2.0**40
        moduli = coeff_modulus_create(n, (60, 40, 40, 60))
        ctx = hg.Context(n, moduli)
        stream = torch.cuda.ExternalStream(ctx.stream)
        cl = Client(ctx, seed=3)
        rk = cl.relin_key()
        ctx.load_relin_key(rk)
        steps = [s * (1 << k) for k in range(7) for s in (1, -1)]
        gk = cl.galois_keys_for_steps(steps)
        ctx.load_galois_keys(gk)
        rng = np.random.default_rng(1)
        o = orc.Oracle(n, moduli)
        if a.only in ("", "bmatmul"):
            th, ot = residues(rng, moduli[:L], (dim, 2), n), residues(rng, moduli[:L], (dim, 2), n)
            T, O, out = ctx.upload_ct(th, sc), ctx.upload_ct(ot, sc), ctx.ct(dim, 2)
            fn = lambda: ctx.bmatmul(out, T, O, dim, dim, True)  # noqa: E731
            ms = timed(torch, ctx, stream, fn, 5)
            ks = sum(len(naf(i)) for i in range(dim)) * dim + dim  # NAF rotations of every (i, j) + one relinearisation per output
            rec = {"config": "cfg3(i): BatchedMatrix::matmul case B, 64x64 encrypted x encrypted, N=16384 {60,40,40,60}, default power-of-two Galois keys (NAF)",
                   "ms_per_matmul": ms, "matmuls_per_s": 1e3 / ms, "key_switches_per_matmul": ks, "key_switches_per_s": ks / (ms * 1e-3),
                   "ct_ct_multiplies_per_matmul": dim * dim, **kernel_split(ctx, fn)}
            if not a.no_cpu:
                def one(i):  # the reference's loop for output i (he_linalg.cpp:977-1003), case B
                    acc = None
                    for j in range(dim):
                        t = o.multiply(o.rotate(ot[j], i, gk)[0] if i else ot[j], th[j])
                        acc = t if acc is None else o.add(acc, t)
                    return o.rescale(o.relinearize(acc, rk))
                sample = list(range(1, 1 + threads))
                t0 = time.perf_counter()
                with ThreadPoolExecutor(threads) as ex:
                    res = list(ex.map(one, sample))
                dt = time.perf_counter() - t0
                ks_s = sum(len(naf(i)) for i in sample) * dim + len(sample)
                got = out.download()
                rec["cpu_port"] = {"cores": threads, "sample": f"outputs 1..{threads} of 64 (one per host thread)", "key_switches_per_s": ks_s / dt,
                                   "matmuls_per_s_extrapolated": (ks_s / dt) / ks, "bit_exact_vs_gpu": bool(all(np.array_equal(got[i], r) for i, r in zip(sample, res)))}
                rec["speedup_vs_cpu_port"] = rec["key_switches_per_s"] / rec["cpu_port"]["key_switches_per_s"]
            emit(rec)
            del T, O, out
        if a.only in ("", "matmul"):
            A = ctx.ct(dim * dim, 2, L)
            Bm = ctx.ct(dim * dim, 2, L)
            A.set_meta(2, L, sc)
            Bm.set_meta(2, L, sc)
            base = residues(rng, moduli[:L], (dim, 2), n)
            for i in range(dim * dim):  # 64 distinct ciphertexts tiled over the 4096 entries (timing only)
                A.upload_one(i, base[i % dim])
                Bm.upload_one(i, base[(i * 7 + 3) % dim])
            out = ctx.ct(dim * dim, 2, L - 1)
            fn = lambda: ctx.matmul_elemwise(out, A, Bm, dim, dim, dim)  # noqa: E731
            ms = timed(torch, ctx, stream, fn, 2)
            rec = {"config": "cfg3(ii): Matrix::matmul 64x64, one ciphertext per entry (2 x 4096 inputs), N=16384 {60,40,40,60}",
                   "ms_per_matmul": ms, "matmuls_per_s": 1e3 / ms, "ct_ct_multiplies_per_matmul": dim**3, "ct_ct_multiplies_per_s": dim**3 / (ms * 1e-3),
                   "key_switches_per_matmul": dim * dim, **kernel_split(ctx, fn)}
            if not a.no_cpu:
                def entry(idx):
                    i, j = idx % dim, idx // dim
                    acc = None
                    for k in range(dim):
                        t = o.multiply(base[(i + k * dim) % dim], base[((k + j * dim) * 7 + 3) % dim])
                        acc = t if acc is None else o.add(acc, t)
                    return o.rescale(o.relinearize(acc, rk))
                sample = list(range(threads))
                t0 = time.perf_counter()
                with ThreadPoolExecutor(threads) as ex:
                    res = list(ex.map(entry, sample))
                dt = time.perf_counter() - t0
                got = np.stack([out.download_one(i) for i in sample])
                rec["cpu_port"] = {"cores": threads, "sample": f"entries 0..{threads - 1} of 4096", "ct_ct_multiplies_per_s": len(sample) * dim / dt,
                                   "matmuls_per_s_extrapolated": len(sample) / dt / (dim * dim), "bit_exact_vs_gpu": bool(all(np.array_equal(g, r) for g, r in zip(got, res)))}
                rec["speedup_vs_cpu_port"] = rec["matmuls_per_s"] / rec["cpu_port"]["matmuls_per_s_extrapolated"]
            emit(rec)
            del A, Bm, out
        del ctx

    # ------------------------------------------------------------ cfg 4 and the shipped bfft
    for name, m, bits, scale in (("bfft4096", 4096, (60,) + (40,) * 12 + (60,), 2.0**40), ("bfft128", 128, (60, 31) + (30,) * 9 + (60,), 2.0**30)):
        if a.only not in ("", name):
            continue
        n = 16384
        moduli = coeff_modulus_create(n, bits)
        K = len(moduli)
        L0 = K - 1
        stages = m.bit_length() - 1
        ctx = hg.Context(n, moduli)
        stream = torch.cuda.ExternalStream(ctx.stream)
        cl = Client(ctx, seed=4)
        steps = [s * (m >> i) for i in range(1, stages + 1) for s in (1, -1)]
        gk = cl.galois_keys_for_steps(steps)
        ctx.load_galois_keys(gk)
        rng = np.random.default_rng(2)
        o = orc.Oracle(n, moduli)
        pts_host = [residues(rng, moduli[:L0 - i], (3,), n) for i in range(stages)]  # D0..D2 of stage i+1 at its level
        pts = [ctx.upload_pt(p_, scale) for p_ in pts_host]
        for B in (1, 32):
            x = residues(rng, moduli[:L0], (B, 2), n)
            X0, Y = ctx.upload_ct(x, scale, size_cap=2, L_cap=L0), ctx.ct(B, 2, L0)

            def fft():
                Y.copy_from(X0)
                for i in range(1, stages + 1):
                    Y.scale = scale  # the reference re-encodes the stage diagonals at y's current scale; random diagonals here
                    ctx.bfft_stage(Y, pts[i - 1], m >> i, i != 1)

            ms = timed(torch, ctx, stream, fft, 3)
            rot = 2 * stages - 1
            rec = {"config": f"{name}: he::fft::bfft over {m} slots, N=16384, chain {list(bits)[:3]}..x{K} primes, {stages} stages, {rot} rotations per FFT",
                   "batch": B, "ms_per_batch": ms, "ffts_per_s": B / (ms * 1e-3), "key_switches_per_s": B * rot / (ms * 1e-3), **kernel_split(ctx, fft)}
            if not a.no_cpu and B == 1:
                def cpu_fft(seed):
                    y = x[0]
                    for i in range(1, stages + 1):
                        st, d = m >> i, pts_host[i - 1]
                        r = o.add(o.rescale(o.multiply_plain(y, d[0])), o.rescale(o.multiply_plain(o.rotate(y, st, gk)[0], d[1])))
                        if i != 1:
                            r = o.add(r, o.rescale(o.multiply_plain(o.rotate(y, -st, gk)[0], d[2])))
                        y = r
                    return y
                t0 = time.perf_counter()
                with ThreadPoolExecutor(threads) as ex:
                    res = list(ex.map(cpu_fft, range(threads)))
                dt = time.perf_counter() - t0
                rec["cpu_port"] = {"cores": threads, "sample": f"{threads} independent FFTs of the same ciphertext, one per host thread", "ffts_per_s": threads / dt,
                                   "key_switches_per_s": threads * rot / dt, "bit_exact_vs_gpu": bool(np.array_equal(res[0], Y.download()[0]))}
            emit(rec)
            del X0, Y
        del ctx, pts


if __name__ == "__main__":
    main()
