#!/usr/bin/env python
"""Generates tests/golden/golden_v1.json and golden_small.npz from the CPU oracle.

PARITY UNPINNED: the reference ships no golden vectors and SEAL cannot be run here, so these
fixtures pin the ORACLE (and through it the CUDA path) against regressions; what ties the oracle
to SEAL's published behaviour is listed in DESIGN.md section 2.

* golden_small.npz -- explicit toy vectors (N = 64): moduli, roots, an NTT pair, a key-switch
  and a rescale input/output (also reproduced by the big-integer restatement in the tests).
* golden_v1.json  -- SHA-256 digests of evaluator outputs at N = 4096 / 8192 on inputs that the
  oracle's own deterministic generators (splitmix64 seeds) rebuild on any machine.

Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def seeded_case(n, bits):
    """Deterministic inputs from the oracle's own generators."""
    moduli = orc.coeff_modulus_create(n, bits)
    o = orc.Oracle(n, moduli)
    L = len(moduli) - 1
    s = o.sample_secret(1)
    rk = o.gen_relin_key(2, s)
    steps = [1, 2, 4, -1]
    gk = {orc.galois_elt_from_step(n, st): o.gen_galois_key(10 + i, s, orc.galois_elt_from_step(n, st)) for i, st in enumerate(steps)}
    zero = np.zeros((L, n), dtype=np.uint64)
    a = o.encrypt_symmetric(20, s, zero)
    b = o.encrypt_symmetric(21, s, zero)
    pt = o.encrypt_symmetric(22, s, zero)[0]  # pseudo-random "plaintext" limbs
    return moduli, o, L, s, rk, gk, a, b, pt


def outputs(n, bits):
    moduli, o, L, s, rk, gk, a, b, pt = seeded_case(n, bits)
    prod = o.multiply(a, b)
    rel = o.relinearize(prod, rk)
    res = {
        "moduli": [hex(q) for q in moduli],
        "psi": [hex(o.psi(i)) for i in range(len(moduli))],
        "in_a": digest(a), "in_b": digest(b), "in_pt": digest(pt), "relin_key": digest(rk),
        "add": digest(o.add(a, b)), "sub": digest(o.sub(a, b)), "negate": digest(o.negate(a)),
        "multiply_plain": digest(o.multiply_plain(a, pt)), "add_plain": digest(o.add_plain(a, pt)),
        "multiply": digest(prod), "square": digest(o.square(a)), "relinearize": digest(rel),
        "rescale": digest(o.rescale(rel)), "mod_switch": digest(o.mod_switch(a)),
        "rotate_1": digest(o.rotate(a, 1, gk)[0]), "rotate_m1": digest(o.rotate(a, -1, gk)[0]),
        "rotate_7_naf": digest(o.rotate(a, 7, {k: v for k, v in gk.items()} | {orc.galois_elt_from_step(n, 8): o.gen_galois_key(30, s, orc.galois_elt_from_step(n, 8))})[0]),
        "ntt_fwd_limb0_of_a": digest(o.ntt_fwd(0, a[0, 0])), "ntt_inv_limb0_of_a": digest(o.ntt_inv(0, a[0, 0])),
    }
    # BSGS matvec 2x2 giant/baby on two ciphertexts, all four modes
    cts = np.stack([a, b])
    pts = np.stack([o.encrypt_symmetric(40 + i, s, np.zeros((L, n), dtype=np.uint64))[0] for i in range(4)])
    bk = [None, gk[orc.galois_elt_from_step(n, 1)]]
    gkeys = [None, gk[orc.galois_elt_from_step(n, 2)]]
    for hoist in (0, 1):
        for lazy in (0, 1):
            res[f"matvec_2x2_hoist{hoist}_lazy{lazy}"] = digest(o.matvec_bsgs(cts, 2, 2, pts, bk, gkeys, hoist=bool(hoist), lazy=bool(lazy)))
    # double-hoisted mode: plaintexts with a limb mod the special prime (pseudo-random residues over all K limbs)
    res["matvec_2x2_dh"] = digest(o.matvec_bsgs(cts, 2, 2, dh_plaintexts(o, s, L, n), bk, gkeys, dh=True))
    # 2 baby x 3 giant steps WITHOUT the final rescale: LAZY (one mod-down for the sum of the two rotated giant
    # steps) differs from the step-by-step form by the rounding of a mod-down, +-1 per coefficient, which the
    # rescale of the 2x2 cases above absorbs (lazy0 == lazy1 there); before the rescale it is visible
    pts3, gkeys3 = matvec_2x3_inputs(o, s, L, n, gk)
    for hoist in (0, 1):
        for lazy in (0, 1):
            res[f"matvec_2x3_hoist{hoist}_lazy{lazy}"] = digest(o.matvec_bsgs(cts, 2, 3, pts3, bk, gkeys3, hoist=bool(hoist), lazy=bool(lazy), rescale=False))
    res["matvec_2x3_dh"] = digest(o.matvec_bsgs(cts, 2, 3, dh_plaintexts(o, s, L, n, 6), bk, gkeys3, dh=True, rescale=False))
    return res


def matvec_2x3_inputs(o, s, L, n, gk):
    pts3 = np.stack([o.encrypt_symmetric(60 + i, s, np.zeros((L, n), dtype=np.uint64))[0] for i in range(6)])
    gkeys3 = [None, gk[orc.galois_elt_from_step(n, 2)], gk[orc.galois_elt_from_step(n, 4)]]
    return pts3, gkeys3


def dh_plaintexts(o, s, L, n, count=4):
    return np.stack([o.encrypt_symmetric(50 + i, s, np.zeros((L + 1, n), dtype=np.uint64))[0] for i in range(count)])


def small_vectors():
    n = 64
    moduli = orc.coeff_modulus_create(n, [30, 25, 28, 30])
    o = orc.Oracle(n, moduli)
    s = o.sample_secret(5)
    rk = o.gen_relin_key(6, s)
    L = 3
    zero = np.zeros((L, n), dtype=np.uint64)
    a = o.encrypt_symmetric(7, s, zero)
    b = o.encrypt_symmetric(8, s, zero)
    prod = o.multiply(a, b)
    rel = o.relinearize(prod, rk)
    return dict(n=n, moduli=np.array(moduli, dtype=np.uint64), psi=np.array([o.psi(i) for i in range(4)], dtype=np.uint64),
                a=a, b=b, relin_key=rk, ntt_in=a[0, 0], ntt_out=o.ntt_fwd(0, a[0, 0]), product=prod, relinearized=rel,
                rescaled=o.rescale(rel))


def main():
    gold = {"format": 2, "generator": "tests/golden/make_golden.py",
            "cases": {"n4096_36_36_37": outputs(4096, [36, 36, 37]), "n8192_60_40_40_60": outputs(8192, [60, 40, 40, 60])}}
    with open(os.path.join(HERE, "golden_v1.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **small_vectors())
    print("wrote", os.path.join(HERE, "golden_v1.json"), os.path.getsize(os.path.join(HERE, "golden_small.npz")), "bytes npz")


if __name__ == "__main__":
    main()
