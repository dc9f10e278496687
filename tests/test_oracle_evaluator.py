"""Oracle pinning, part 2: evaluator primitives.

(a) switch_key / rescale of the C oracle against an independent big-integer restatement
    of SURVEY.md 9.6 / 9.7 (tests/ckks_ref.py) at toy sizes, bit-exact;
(b) homomorphic correctness after decryption against float64 numpy at the reference's
    parameter sets (matrix_operations.cpp:1048-1052)."""
import numpy as np
import pytest

import ckks_ref as ref
from fixtures import ckks_tol
from oracle import oracle as orc


def _rand_poly(rng, moduli, shape_prefix, n):
    out = np.empty(tuple(shape_prefix) + (len(moduli), n), dtype=np.uint64)
    for i, q in enumerate(moduli):
        out[..., i, :] = rng.integers(0, q, size=tuple(shape_prefix) + (n,), dtype=np.uint64)
    return out


@pytest.mark.parametrize("bits,L", [([30, 25, 28, 30], 3), ([30, 25, 28, 30], 2), ([20, 30, 25], 2), ([30, 30], 1)])
def test_switch_key_and_rescale_vs_bigint(bits, L):
    n = 16
    moduli = orc.coeff_modulus_create(n, bits)
    K = len(moduli)
    o = orc.Oracle(n, moduli)
    psis = [o.psi(i) for i in range(K)]
    rng = np.random.default_rng(7)
    ct = _rand_poly(rng, moduli[:L], (2,), n)
    target = _rand_poly(rng, moduli[:L], (), n)
    key = _rand_poly(rng, moduli, (K - 1, 2), n)
    got = o.switch_key(ct, target, key)
    want = ref.switch_key_ref(ct.tolist(), target.tolist(), key.tolist(), moduli, psis, L)
    assert got.tolist() == want
    if L >= 2:
        ct3 = _rand_poly(rng, moduli[:L], (3,), n)
        assert o.rescale(ct3).tolist() == ref.rescale_ref(ct3.tolist(), moduli, psis, L)
        assert np.array_equal(o.mod_switch(ct3), ct3[:, : L - 1, :])


def test_fast_bigint_transforms_match_direct_evaluation():
    """ntt_fast / intt_fast (O(N log N), used for the full-size cross-check below) are the same functions as the
    O(N^2) direct evaluation of SURVEY 9.2."""
    rng = np.random.default_rng(11)
    for n, bits in ((64, 30), (128, 41), (32, 60)):
        q = ref.get_primes(2 * n, bits, 1)[0]
        psi = ref.minimal_primitive_root(q, n)
        a = [int(x) for x in rng.integers(0, q, n, dtype=np.uint64)]
        f = ref.ntt_fast(a, q, psi)
        assert f == ref.ntt_naive(a, q, psi)
        assert ref.intt_fast(f, q, psi) == a == ref.intt_naive(f, q, psi)


@pytest.mark.parametrize("n,L", [(4096, 3), (8192, 2)])
def test_switch_key_rescale_rotate_vs_bigint_full_size(n, L):
    """the independent big-integer restatement at the reference's own chain {60,40,40,60}
    (matrix_operations.cpp:1050) and at full ring degrees, not only N = 64: switch_key (= relinearize),
    rescale and rotate_vector (key present, and a NAF chain 3 = [-1, 4]) of the C oracle, bit for bit."""
    moduli = orc.coeff_modulus_create(n, [60, 40, 40, 60])
    K = len(moduli)
    o = orc.Oracle(n, moduli)
    psis = [o.psi(i) for i in range(K)]
    assert psis == [ref.minimal_primitive_root(q, n) for q in moduli]
    rng = np.random.default_rng(70 + L)
    ct = _rand_poly(rng, moduli[:L], (2,), n)
    target = _rand_poly(rng, moduli[:L], (), n)
    key = _rand_poly(rng, moduli, (K - 1, 2), n)
    s = o.sample_secret(3)
    steps = [-1, 4]
    gk = {orc.galois_elt_from_step(n, st): o.gen_galois_key(40 + i, s, orc.galois_elt_from_step(n, st)) for i, st in enumerate(steps)}
    with ref.fast_transforms():
        assert o.switch_key(ct, target, key).tolist() == ref.switch_key_ref(ct.tolist(), target.tolist(), key.tolist(), moduli, psis, L)
        ct3 = _rand_poly(rng, moduli[:L], (3,), n)
        assert o.rescale(ct3).tolist() == ref.rescale_ref(ct3.tolist(), moduli, psis, L)
        # rotate by 3 without a key for 3: SEAL's NAF order, LSB first: -1, then 4
        got, nks = o.rotate(ct, 3, gk)
        assert nks == 2
        want = ct.tolist()
        for st in steps:
            elt = ref.galois_elt_from_step(n, st)
            assert elt == orc.galois_elt_from_step(n, st)
            want = ref.apply_galois_ref(want, elt, gk[elt].tolist(), moduli, psis, L)
        assert got.tolist() == want


@pytest.fixture(scope="module")
def setup8192():
    n = 8192
    moduli = orc.coeff_modulus_create(n, [60, 40, 40, 60])
    o = orc.Oracle(n, moduli)
    enc = ref.Encoder(n, moduli, o.ntt_fwd, o.ntt_inv)
    s = o.sample_secret(1234)
    return n, moduli, o, enc, s


def test_encrypt_decrypt_roundtrip(setup8192):
    n, moduli, o, enc, s = setup8192
    rng = np.random.default_rng(0)
    z = rng.uniform(-1, 1, n // 2) + 1j * rng.uniform(-1, 1, n // 2)
    scale = 2.0**40
    pt = enc.encode(z, scale, 3)
    assert np.max(np.abs(enc.decode(pt, scale) - z)) < 1e-9
    ct = o.encrypt_symmetric(99, s, pt)
    dec = enc.decode(o.decrypt(ct, s), scale)
    assert np.max(np.abs(dec - z)) < 1e-7


def test_homomorphic_ops(setup8192):
    n, moduli, o, enc, s = setup8192
    rng = np.random.default_rng(1)
    slots = n // 2
    scale = 2.0**40
    x = rng.uniform(-1, 1, slots)
    y = rng.uniform(-1, 1, slots)
    ptx, pty = enc.encode(x, scale, 3), enc.encode(y, scale, 3)
    cx, cy = o.encrypt_symmetric(1, s, ptx), o.encrypt_symmetric(2, s, pty)
    rk = o.gen_relin_key(3, s)

    def dec(ct, sc):
        return enc.decode(o.decrypt(ct, s), sc).real

    assert np.max(np.abs(dec(o.add(cx, cy), scale) - (x + y))) < 1e-7
    assert np.max(np.abs(dec(o.sub(cx, cy), scale) - (x - y))) < 1e-7
    assert np.max(np.abs(dec(o.negate(cx), scale) + x)) < 1e-7
    assert np.max(np.abs(dec(o.add_plain(cx, pty), scale) - (x + y))) < 1e-7
    assert np.max(np.abs(dec(o.sub_plain(cx, pty), scale) - (x - y))) < 1e-7
    # multiply_plain + rescale
    mp = o.rescale(o.multiply_plain(cx, pty))
    assert np.max(np.abs(dec(mp, scale * scale / moduli[2]) - x * y)) < 1e-6
    # multiply (size 3 decrypts), relinearize, rescale
    m3 = o.multiply(cx, cy)
    assert m3.shape[0] == 3
    assert np.max(np.abs(dec(m3, scale * scale) - x * y)) < 1e-6
    m2 = o.relinearize(m3, rk)
    assert np.max(np.abs(dec(m2, scale * scale) - x * y)) < 1e-6
    r = o.rescale(m2)
    assert np.max(np.abs(dec(r, scale * scale / moduli[2]) - x * y)) < 1e-6
    # square == multiply(a, a); size-3 add/sub semantics
    assert np.array_equal(o.square(cx), o.multiply(cx, cx))
    a23 = o.add(cx, m3)
    assert np.array_equal(a23[2], m3[2])
    s23 = o.sub(cx, m3)
    assert np.array_equal(s23[2], o.negate(m3)[2])


def test_rotation_and_naf(setup8192):
    n, moduli, o, enc, s = setup8192
    rng = np.random.default_rng(2)
    slots = n // 2
    scale = 2.0**40
    x = rng.uniform(-1, 1, slots)
    cx = o.encrypt_symmetric(5, s, enc.encode(x, scale, 3))
    gk = o.gen_galois_keys_for_steps(100, s, [1, 2, 4, 8, -1, -2, 16])

    def dec(ct):
        return enc.decode(o.decrypt(ct, s), scale).real

    r1, nks = o.rotate(cx, 1, gk)
    assert nks == 1
    assert np.max(np.abs(dec(r1) - np.roll(x, -1))) < 1e-6
    r7, nks = o.rotate(cx, 7, gk)  # NAF 7 = [-1, 8]
    assert nks == 2
    assert np.max(np.abs(dec(r7) - np.roll(x, -7))) < 1e-6
    # NAF order is part of the contract: rotate(7) == rotate(rotate(ct,-1), 8)
    step_a, _ = o.rotate(cx, -1, gk)
    step_b, _ = o.rotate(step_a, 8, gk)
    assert np.array_equal(r7, step_b)
    r0, nks = o.rotate(cx, 0, gk)
    assert nks == 0 and np.array_equal(r0, cx)
    with pytest.raises(ValueError):
        o.rotate(cx, 32, gk)  # single-term NAF without a key
    # lower level (after a rescale) uses the same keys
    low = o.rescale(o.multiply_plain(cx, enc.encode_scalar(1.0, scale, 3)))
    r, _ = o.rotate(low, 2, gk)
    d = enc.decode(o.decrypt(r, s), scale * scale / moduli[2]).real
    assert np.max(np.abs(d - np.roll(x, -2))) < 1e-6


def test_matvec_bsgs_decrypts_to_matvec():
    n = 8192
    moduli = orc.coeff_modulus_create(n, [60, 40, 40, 60])
    o = orc.Oracle(n, moduli)
    enc = ref.Encoder(n, moduli, o.ntt_fwd, o.ntt_inv)
    s = o.sample_secret(77)
    rng = np.random.default_rng(3)
    dim, n1, n2 = 16, 4, 4
    slots = n // 2
    scale = 2.0**40
    M = rng.uniform(-1, 1, (dim, dim))
    v = rng.uniform(-1, 1, dim)
    vrep = np.tile(v, slots // dim)
    ct = o.encrypt_symmetric(8, s, enc.encode(vrep, scale, 3))
    pts = np.empty((dim, 3, n), dtype=np.uint64)
    for g in range(n2):
        for b in range(n1):
            d = g * n1 + b
            diag = np.array([M[r, (r + d) % dim] for r in range(dim)])
            pts[d] = enc.encode(np.roll(np.tile(diag, slots // dim), g * n1), scale, 3)
    bk = [None] + [o.gen_galois_key(200 + b, s, orc.galois_elt_from_step(n, b)) for b in range(1, n1)]
    gkeys = [None] + [o.gen_galois_key(300 + g, s, orc.galois_elt_from_step(n, g * n1)) for g in range(1, n2)]
    outs = []
    for fast in (False, True):
        out = o.matvec_bsgs(ct[None], n1, n2, pts, bk, gkeys, threads=2, fast=fast)
        got = enc.decode(o.decrypt(out[0], s), scale * scale / moduli[2]).real[:dim]
        assert np.max(np.abs(got - M @ v)) < ckks_tol(dim, n, scale, "hoisted" if fast else "exact")
        outs.append(out)
    # hoisting changes the digit representatives: same plaintext, different ciphertext bits
    assert not np.array_equal(outs[0], outs[1])


def test_key_switch_phases_compose_to_switch_key():
    """The hoisted restatement is built from the same three phases as orc_switch_key; with one giant
    step and no baby rotation the fast matvec must equal the exact one bit-for-bit."""
    n = 4096
    moduli = orc.coeff_modulus_create(n, [36, 36, 37])
    o = orc.Oracle(n, moduli)
    rng = np.random.default_rng(4)
    s = o.sample_secret(5)
    cts = _rand_poly(rng, moduli[:2], (2, 2), n)
    pts = _rand_poly(rng, moduli[:2], (2,), n)
    gk = o.gen_galois_key(9, s, orc.galois_elt_from_step(n, 1))
    # n1 = 1, n2 = 2: a single giant rotation, the lazy mod-down degenerates to one ordinary key-switch
    a = o.matvec_bsgs(cts, 1, 2, pts, [None], [None, gk], fast=False)
    b = o.matvec_bsgs(cts, 1, 2, pts, [None], [None, gk], fast=True)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("n1,n2,g_split", [(4, 4, None), (8, 2, None), (16, 1, None), (1, 16, None), (4, 4, 2)])
def test_matvec_bsgs_double_hoisted_decrypts(n1, n2, g_split):
    """HEGPU_MATVEC_DH restatement (baby rotations kept in the extended basis, plaintexts with a limb
    mod P, one mod-down per giant step): decrypts to M @ v within the CKKS tolerance, and -- sharded
    by giant steps without the rescale -- the partial ciphertexts sum to the unsharded one up to the
    rounding of the separate mod-downs."""
    n = 8192
    moduli = orc.coeff_modulus_create(n, [60, 40, 40, 60])
    o = orc.Oracle(n, moduli)
    enc = ref.Encoder(n, moduli, o.ntt_fwd, o.ntt_inv)
    s = o.sample_secret(78)
    rng = np.random.default_rng(5)
    dim, L = 16, 3
    slots = n // 2
    scale = 2.0**40
    M = rng.uniform(-1, 1, (dim, dim))
    v = rng.uniform(-1, 1, dim)
    ct = o.encrypt_symmetric(9, s, enc.encode(np.tile(v, slots // dim), scale, L))
    ptsx = np.empty((dim, L + 1, n), dtype=np.uint64)
    for g in range(n2):
        for b in range(n1):
            d = g * n1 + b
            diag = np.array([M[r, (r + d) % dim] for r in range(dim)])
            ptsx[d] = enc.encode_ext(np.roll(np.tile(diag, slots // dim), g * n1), scale, L)
    bk = [None] + [o.gen_galois_key(200 + b, s, orc.galois_elt_from_step(n, b)) for b in range(1, n1)]
    gkeys = [None] + [o.gen_galois_key(300 + g, s, orc.galois_elt_from_step(n, g * n1)) for g in range(1, n2)]
    tol = ckks_tol(dim, n, scale, "hoisted")
    out = o.matvec_bsgs(ct[None], n1, n2, ptsx, bk, gkeys, threads=2, dh=True)
    got = enc.decode(o.decrypt(out[0], s), scale * scale / moduli[2]).real[:dim]
    assert np.max(np.abs(got - M @ v)) < tol
    # the data limbs of the extended plaintexts are the ordinary plaintexts
    assert np.array_equal(ptsx[0, :L], enc.encode(np.tile(np.array([M[r, r % dim] for r in range(dim)]), slots // dim), scale, L))
    if g_split:
        parts = []
        for g0, cnt in ((0, g_split), (g_split, n2 - g_split)):
            parts.append(o.matvec_bsgs(ct[None], n1, cnt, ptsx[g0 * n1:(g0 + cnt) * n1], bk, gkeys[g0:g0 + cnt], threads=2,
                                       dh=True, rescale=False, g_first=g0)[0])
        summed = o.add(parts[0], parts[1])
        got2 = enc.decode(o.decrypt(o.rescale(summed), s), scale * scale / moduli[2]).real[:dim]
        assert np.max(np.abs(got2 - M @ v)) < tol


@pytest.mark.parametrize("bits,L,n1,n2,g_first", [([30, 25, 28, 30], 3, 2, 2, 0), ([30, 25, 28, 30], 2, 3, 2, 0),
                                                    ([20, 30, 25], 2, 2, 3, 1), ([30, 30], 1, 4, 1, 0)])
def test_matvec_double_hoisted_vs_bigint(bits, L, n1, n2, g_first):
    """The C restatement of the double-hoisted matvec against an independent big-integer one, bit-exact."""
    n = 16
    moduli = orc.coeff_modulus_create(n, bits)
    K = len(moduli)
    o = orc.Oracle(n, moduli)
    psis = [o.psi(i) for i in range(K)]
    rng = np.random.default_rng(11)
    ext_mods = moduli[:L] + [moduli[K - 1]]
    ct = _rand_poly(rng, moduli[:L], (2,), n)
    ptsx = _rand_poly(rng, ext_mods, (n1 * n2,), n)
    bk = [None] + [_rand_poly(rng, moduli, (K - 1, 2), n) for _ in range(1, n1)]
    gkeys = [_rand_poly(rng, moduli, (K - 1, 2), n) for _ in range(n2)]
    rescale = L >= 2
    got = o.matvec_bsgs(ct[None], n1, n2, ptsx, bk, gkeys, dh=True, rescale=rescale, g_first=g_first)[0]
    want = ref.matvec_dh_ref(ct.tolist(), n1, n2, ptsx.tolist(), [None if k is None else k.tolist() for k in bk],
                             [k.tolist() for k in gkeys], moduli, psis, L, rescale=rescale, g_first=g_first)
    assert got.tolist() == want


def test_mod_down_commutes_with_galois():
    """The mod-down by the (odd) special prime commutes with a Galois map bit for bit: the map acts limb-wise, also
    on the limb mod P, and the centred remainder is symmetric.  The double-hoisted matvec relies on it twice:
    component 0 of a rotated giant step is permuted in the extended basis, and the rotation of component 1 may be
    applied before or after its mod-down."""
    import random

    n, L = 64, 2
    R = ref
    mods = R.coeff_modulus_create(n, [30, 30, 31])
    psis = [R.minimal_primitive_root(q, n) for q in mods]
    rnd = random.Random(3)
    for step in (1, 3, -2):
        tab = R.galois_table_ntt(n, R.galois_elt_from_step(n, step))
        acc = [[rnd.randrange(m) for _ in range(n)] for m in mods[:L] + [mods[-1]]]
        down_then_rot = [[row[tab[x]] for x in range(n)] for row in R._mod_down_one_ref([list(r) for r in acc], mods, psis, L)]
        rot_then_down = R._mod_down_one_ref([[row[tab[x]] for x in range(n)] for row in acc], mods, psis, L)
        assert down_then_rot == rot_then_down
