#!/usr/bin/env python
"""NTT micro-benchmark (BASELINE.md config "mu"): forward + inverse negacyclic NTT over a
batch of limb polynomials resident in HBM, GB/s at the algorithmic 16*N bytes per limb.

  python tools/ntt_bench.py [--n 16384] [--count 4096] [--iters 20] [--bits 60,40,40,60]
  python tools/ntt_bench.py --sweep      # N in {8192, 16384, 32768} x batches of 1 .. 4096 limbs, one JSON line each
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--count", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--bits", default="60,40,40,60")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--sweep", action="store_true")
    a = ap.parse_args()
    if a.sweep:
        for n in (8192, 16384, 32768):
            for count in (1, 16, 256, 4096):
                a.n, a.count, a.check = n, count, count == 16
                run(a)
    else:
        run(a)


def run(a):
    import torch

    import hegpu_loader

    hg = hegpu_loader.load()
    from hegpu_b200.client import coeff_modulus_create

    bits = [int(x) for x in a.bits.split(",")]
    moduli = coeff_modulus_create(a.n, bits)
    ctx = hg.Context(a.n, moduli)
    K = len(moduli)
    rng = np.random.default_rng(0)
    host = np.empty((a.count, a.n), dtype=np.uint64)
    for i in range(a.count):
        host[i] = rng.integers(0, moduli[i % K], size=a.n, dtype=np.uint64)
    d = torch.from_numpy(host.view(np.int64)).cuda()
    stream = torch.cuda.ExternalStream(ctx.stream)
    torch.cuda.synchronize()
    out = {"n": a.n, "count": a.count, "bits": bits, "bytes_per_limb": 16 * a.n}
    peak = None
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    for name, fn in (("fwd", ctx.ntt_forward_device), ("inv", ctx.ntt_inverse_device)):
        for _ in range(3):
            fn(d.data_ptr(), a.count, 0, K)
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.iters):
            fn(d.data_ptr(), a.count, 0, K)
        e1.record(stream)
        ctx.sync()
        ms = e0.elapsed_time(e1) / a.iters
        gbs = a.count * 16 * a.n / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "GBps": gbs, "frac_of_measured_peak": gbs / peak if peak else None}
    if a.check:  # fwd then inv the same number of times must return the input
        back = d.cpu().numpy().view(np.uint64)
        out["roundtrip_ok"] = bool(np.array_equal(back, host))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
