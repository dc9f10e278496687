"""Oracle pinning, part 1: parameter generation, NTT conventions, Galois tables, NAF.

Known answers: the prime chains of SURVEY.md 9.1 (derived from SEAL's get_primes /
CoeffModulus::Create rule) and the chain SEAL's own `4_levels` example prints for
CoeffModulus::Create(8192, {50,30,30,50,50}) [recollected from SEAL's examples]."""
import numpy as np
import pytest

import ckks_ref as ref
from oracle import oracle as orc

CHAINS = [
    (8192, [60, 40, 40, 60], [0xFFFFFFFFFFE8001, 0xFFFFF4C001, 0xFFFFFDC001, 0xFFFFFFFFFFFC001]),
    (16384, [60] + [30] * 10 + [60],
     [0xFFFFFFFFFFD8001, 0x3FD20001, 0x3FD78001, 0x3FDC8001, 0x3FDE0001, 0x3FED0001, 0x3FF28001, 0x3FF58001,
      0x3FF78001, 0x3FFC0001, 0x3FFE8001, 0xFFFFFFFFFFE8001]),
    (32768, [60] + [40] * 5 + [60],
     [0xFFFFFFFFF840001, 0xFFFF8A0001, 0xFFFF940001, 0xFFFFB20001, 0xFFFFC40001, 0xFFFFE80001, 0xFFFFFFFFFFC0001]),
    # SEAL examples/4_levels.cpp prints this chain
    (8192, [50, 30, 30, 50, 50], [0x3FFFFFFEF4001, 0x3FFE8001, 0x3FFF4001, 0x3FFFFFFFCC001, 0x3FFFFFFFFC001]),
]


@pytest.mark.parametrize("n,bits,expect", CHAINS)
def test_prime_chain_known_answers(n, bits, expect):
    assert orc.coeff_modulus_create(n, bits) == expect
    assert ref.coeff_modulus_create(n, bits) == expect


def test_fft_demo_chain_31_bit():
    # fft.cpp:135-140 {60,31,30x9,60}
    got = orc.coeff_modulus_create(16384, [60, 31] + [30] * 9 + [60])
    assert got[1] == 0x7FFE0001
    assert got[2] == 0x3FD78001 and got[-1] == 0xFFFFFFFFFFE8001
    assert got == ref.coeff_modulus_create(16384, [60, 31] + [30] * 9 + [60])


def test_minimal_root_and_monomial():
    n = 64
    moduli = orc.coeff_modulus_create(n, [30, 25, 30])
    o = orc.Oracle(n, moduli)
    for i, q in enumerate(moduli):
        psi = o.psi(i)
        assert psi == ref.minimal_primitive_root(q, n)
        assert pow(psi, n, q) == q - 1
        x = np.zeros(n, dtype=np.uint64)
        x[1] = 1
        got = o.ntt_fwd(i, x)
        want = [pow(psi, 2 * ref.brev(k, 6) + 1, q) for k in range(n)]
        assert got.tolist() == want


@pytest.mark.parametrize("n", [16, 64, 256])
def test_ntt_matches_direct_evaluation(n):
    rng = np.random.default_rng(n)
    moduli = orc.coeff_modulus_create(n, [40, 30, 40])
    o = orc.Oracle(n, moduli)
    for i, q in enumerate(moduli):
        a = rng.integers(0, q, size=n, dtype=np.uint64)
        f = o.ntt_fwd(i, a)
        assert f.tolist() == ref.ntt_naive(a.tolist(), q, o.psi(i))
        assert o.ntt_inv(i, f).tolist() == a.tolist()
        assert ref.intt_naive(f.tolist(), q, o.psi(i)) == a.tolist()


@pytest.mark.parametrize("n", [8192, 16384, 32768])
def test_ntt_roundtrip_and_convolution_full_size(n):
    bits = [60, 40, 60]
    moduli = orc.coeff_modulus_create(n, bits)
    o = orc.Oracle(n, moduli)
    rng = np.random.default_rng(1)
    for i, q in enumerate(moduli):
        a = rng.integers(0, q, size=n, dtype=np.uint64)
        assert np.array_equal(o.ntt_inv(i, o.ntt_fwd(i, a)), a)
        # negacyclic: X * a(X) in the NTT domain is a^ (.) NTT(X)
        x = np.zeros(n, dtype=np.uint64)
        x[1] = 1
        xa = np.empty(n, dtype=np.uint64)
        xa[1:] = a[:-1]
        xa[0] = (q - int(a[-1])) % q
        fa, fx = o.ntt_fwd(i, a).tolist(), o.ntt_fwd(i, x).tolist()
        prod = np.array([u * v % q for u, v in zip(fa, fx)], dtype=np.uint64)
        assert np.array_equal(prod, o.ntt_fwd(i, xa))


def test_naf_known_answers():
    # SURVEY 9.4
    assert orc.naf(3) == [-1, 4]
    assert orc.naf(5) == [1, 4]
    assert orc.naf(6) == [-2, 8]
    assert orc.naf(7) == [-1, 8]
    assert orc.naf(11) == [-1, -4, 16]
    assert orc.naf(63) == [-1, 64]
    assert orc.naf(-3) == [1, -4]
    # SEAL's own unit test (native/tests/seal/util/numth.cpp, NAF)
    for v, want in ((0, []), (1, [1]), (-1, [-1]), (2, [2]), (-2, [-2]), (127, [-1, 128]), (-127, [1, -128]), (123, [-1, -4, 128]),
                    (-123, [1, 4, -128])):
        assert orc.naf(v) == want, v
    for v in range(-300, 300):
        assert sum(orc.naf(v)) == v


def test_galois_elt_and_table():
    n = 64
    assert orc.galois_elt_from_step(n, 0) == 2 * n - 1
    assert orc.galois_elt_from_step(n, 1) == 3
    assert orc.galois_elt_from_step(n, -1) == pow(3, n // 2 - 1, 2 * n)
    moduli = orc.coeff_modulus_create(n, [30, 30])
    o = orc.Oracle(n, moduli)
    rng = np.random.default_rng(5)
    q = moduli[0]
    a = rng.integers(0, q, size=n, dtype=np.uint64)
    for step in (1, 2, 5, -3, 31):
        elt = orc.galois_elt_from_step(n, step)
        want = o.ntt_fwd(0, np.array(ref.galois_coeff(a.tolist(), elt, q), dtype=np.uint64))
        got = o.apply_galois_ntt(o.ntt_fwd(0, a), elt)
        assert np.array_equal(got, want)
        # block structure the CUDA gather relies on: aligned 2^k blocks map to aligned blocks
        t = o.galois_table(elt)
        for k in (2, 8, 32):
            assert np.array_equal(t.reshape(-1, k)[:, 0] // k, t.reshape(-1, k)[:, -1] // k)


def test_seal_unit_test_known_answers():
    """Known answers of Microsoft SEAL's OWN unit tests (native/tests/seal/util/{numth,ntt,galois}.cpp, SEAL 4.x), quoted
    from the published test sources -- the only reference-held vectors this path has (SEAL itself is absent).
    They pin the three conventions everything else rests on: which primitive root SEAL picks (the numerically smallest),
    the order and values of the forward transform, and the NTT-form Galois permutation.  The constants are mutually
    consistent (checked below), which a misremembered digit would break."""
    q = 0xFFFFFFFFFFC0001
    # numth.cpp TryMinimalPrimitiveRoot: (modulus, degree) -> root
    for mod, degree, root in ((11, 2, 10), (29, 2, 28), (29, 4, 12), (1234565441, 2, 1234565440), (1234565441, 8, 249725733)):
        assert pow(root, degree // 2, mod) == mod - 1
        assert ref.minimal_primitive_root(mod, degree // 2) == root
        if degree >= 4:
            assert orc.Oracle(degree // 2, [mod]).psi(0) == root
    # ntt.cpp NTTBasics: coeff_count_power 2, modulus 0xffffffffffc0001: root powers in bit-reversed order 1, psi^2, psi, psi^3
    psi, psi2, psi3 = 178930308976060547, 288794978602139552, 748001537669050592
    assert pow(psi, 2, q) == psi2 and pow(psi, 3, q) == psi3 and pow(psi, 4, q) == q - 1
    assert orc.Oracle(4, [q]).psi(0) == psi
    # ntt.cpp NegacyclicNTTTest: coeff_count_power 1: {0,0} -> {0,0}; {1,0} -> {1,1}; {1,1} -> {288794978602139553, 864126526004445282}
    o = orc.Oracle(2, [q])
    assert o.psi(0) == psi2
    for poly, want in (([0, 0], [0, 0]), ([1, 0], [1, 1]), ([1, 1], [288794978602139553, 864126526004445282])):
        got = o.ntt_fwd(0, np.array(poly, dtype=np.uint64))
        assert got.tolist() == want
        assert o.ntt_inv(0, got).tolist() == poly  # InverseNegacyclicNTTTest: round trip
    # galois.cpp: GaloisTool(3) (N = 8): get_elt_from_step and the two automorphism forms on in = {0..7}
    assert [orc.galois_elt_from_step(8, s) for s in (0, 1, -3, 2, -2, 3, -1)] == [15, 3, 3, 9, 9, 11, 11]
    assert ref.galois_coeff(list(range(8)), 3, 17) == [0, 14, 6, 1, 13, 7, 2, 12]      # ApplyGalois, modulus 17
    assert orc.Oracle(8, [q]).galois_table(3).tolist() == [4, 5, 7, 6, 1, 0, 2, 3]       # ApplyGaloisNTT: out[i] = in[table[i]]
