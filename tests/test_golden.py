"""Committed golden fixtures (tests/golden/, made by make_golden.py from the oracle).
CPU: the oracle still reproduces them (and the big-integer restatement agrees on the toy
vectors).  GPU: libhegpu.so hashes to the same digests on the same seeded inputs."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

import ckks_ref as ref
import hegpu_loader
from oracle import oracle as orc

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, HERE)
import make_golden  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden_v1.json")))
CASES = {"n4096_36_36_37": (4096, [36, 36, 37]), "n8192_60_40_40_60": (8192, [60, 40, 40, 60])}


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden_digests(name):
    n, bits = CASES[name]
    assert make_golden.outputs(n, bits) == GOLD["cases"][name]


def test_golden_matvec_digests_are_not_vacuous():
    """Every mode of the 2x3 matvec (taken before the rescale) is a different ciphertext: the modes differ in where
    they round, also the lazy flag alone (the rescaled 2x2 digests hide that +-1)."""
    for name in CASES:
        g = GOLD["cases"][name]
        d = [g[f"matvec_2x3_hoist{h}_lazy{l}"] for h in (0, 1) for l in (0, 1)] + [g["matvec_2x3_dh"]]
        assert len(set(d)) == len(d)


def test_small_vectors_against_oracle_and_bigint():
    z = np.load(os.path.join(HERE, "golden_small.npz"))
    n, moduli = int(z["n"]), [int(q) for q in z["moduli"]]
    assert moduli == orc.coeff_modulus_create(n, [30, 25, 28, 30])
    o = orc.Oracle(n, moduli)
    psis = [int(p) for p in z["psi"]]
    assert psis == [ref.minimal_primitive_root(q, n) for q in moduli]
    assert ref.ntt_naive(z["ntt_in"].tolist(), moduli[0], psis[0]) == z["ntt_out"].tolist()
    assert np.array_equal(o.multiply(z["a"], z["b"]), z["product"])
    prod = z["product"]
    want = ref.switch_key_ref(prod[:2].tolist(), prod[2].tolist(), z["relin_key"].tolist(), moduli, psis, 3)
    assert z["relinearized"].tolist() == want
    assert z["rescaled"].tolist() == ref.rescale_ref(z["relinearized"].tolist(), moduli, psis, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_matches_golden_digests(name):
    hg = hegpu_loader.load()
    n, bits = CASES[name]
    g = GOLD["cases"][name]
    moduli, o, L, s, rk, gk, a, b, pt = make_golden.seeded_case(n, bits)  # inputs only
    assert digest(a) == g["in_a"] and digest(rk) == g["relin_key"]
    ctx = hg.Context(n, moduli)
    assert [hex(ctx.psi(i)) for i in range(len(moduli))] == g["psi"]
    ctx.load_relin_key(rk)
    gk8 = o.gen_galois_key(30, s, orc.galois_elt_from_step(n, 8))
    ctx.load_galois_keys(gk)
    sc = 2.0**20
    A, B = ctx.upload_ct(a, sc), ctx.upload_ct(b, sc)
    P = ctx.upload_pt(pt, sc)
    out = ctx.ct(1)

    def chk(key):
        assert digest(out.download()[0]) == g[key], key

    ctx.add(out, A, B); chk("add")
    ctx.sub(out, A, B); chk("sub")
    ctx.negate(out, A); chk("negate")
    ctx.multiply_plain(out, A, P, 0); chk("multiply_plain")
    ctx.add_plain(out, A, P, 0); chk("add_plain")
    ctx.square(out, A); chk("square")
    ctx.mod_switch_to_next(out, A); chk("mod_switch")
    ctx.rotate_vector(out, A, 1); chk("rotate_1")
    ctx.rotate_vector(out, A, -1); chk("rotate_m1")
    ctx.multiply(out, A, B); chk("multiply")
    ctx.relinearize(out, out); chk("relinearize")
    ctx.rescale_to_next(out, out); chk("rescale")
    ctx.load_galois_key(orc.galois_elt_from_step(n, 8), gk8)
    ctx.rotate_vector(out, A, 7); chk("rotate_7_naf")
    limb = np.ascontiguousarray(a[0, 0:1]).copy()
    ctx.ntt_forward_host(limb, 0, 1)
    assert digest(limb[0]) == g["ntt_fwd_limb0_of_a"]
    limb = np.ascontiguousarray(a[0, 0:1]).copy()
    ctx.ntt_inverse_host(limb, 0, 1)
    assert digest(limb[0]) == g["ntt_inv_limb0_of_a"]
    cts = np.stack([a, b])
    pts = np.stack([o.encrypt_symmetric(40 + i, s, np.zeros((L, n), dtype=np.uint64))[0] for i in range(4)])
    X, D = ctx.upload_ct(cts, sc, size_cap=2), ctx.upload_pt(pts, sc)
    mv = ctx.ct(2, 2)
    for hoist in (0, 1):
        for lazy in (0, 1):
            ctx.matvec_bsgs(mv, X, D, 2, 2, hoist=bool(hoist), lazy=bool(lazy))
            assert digest(mv.download()) == g[f"matvec_2x2_hoist{hoist}_lazy{lazy}"]
    Dx = ctx.upload_pt_ext(make_golden.dh_plaintexts(o, s, L, n), sc)
    ctx.matvec_bsgs(mv, X, Dx, 2, 2, dh=True)
    assert digest(mv.download()) == g["matvec_2x2_dh"]
    pts3, _ = make_golden.matvec_2x3_inputs(o, s, L, n, gk)
    D3 = ctx.upload_pt(pts3, sc)
    for hoist in (0, 1):
        for lazy in (0, 1):
            ctx.matvec_bsgs(mv, X, D3, 2, 3, hoist=bool(hoist), lazy=bool(lazy), rescale=False)
            assert digest(mv.download()) == g[f"matvec_2x3_hoist{hoist}_lazy{lazy}"]
    Dx3 = ctx.upload_pt_ext(make_golden.dh_plaintexts(o, s, L, n, 6), sc)
    ctx.matvec_bsgs(mv, X, Dx3, 2, 3, dh=True, rescale=False)
    assert digest(mv.download()) == g["matvec_2x3_dh"]
