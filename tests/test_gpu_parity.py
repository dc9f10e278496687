"""GPU parity tests proper: every evaluator primitive of libhegpu.so (called through the
C ABI) must be BIT-EXACT against the CPU oracle on the same inputs (integer arithmetic;
tolerance = 0).  Decrypted composites are additionally checked against float64 numpy
within the CKKS noise tolerance fixtures.ckks_tol stated in DESIGN.md."""
import numpy as np
import pytest

import hegpu_loader
from fixtures import ckks_tol, rand_residues, setup
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hg():
    return hegpu_loader.load()


def make_ctx(hg, S):
    return hg.Context(S.n, S.moduli)


# ------------------------------------------------------------------ NTT
@pytest.mark.parametrize("n", [4096, 8192, 16384, 32768])
def test_ntt_bit_exact(hg, n):
    bits = [60, 50, 40, 30, 59, 47]  # covers both lazy-range code paths (fwd: 2^58, inv: 2^46)
    S = setup(n, tuple(bits))
    ctx = make_ctx(hg, S)
    rng = np.random.default_rng(n)
    for i, q in enumerate(S.moduli):
        assert ctx.psi(i) == S.o.psi(i)
    a = rand_residues(rng, S.moduli, (3,), n)  # [3][K][N]
    a[0, :, :] = 0
    for i, q in enumerate(S.moduli):
        a[1, i, :] = q - 1
    f = a.copy()
    ctx.ntt_forward_host(f, 0, S.K)
    want = np.stack([np.stack([S.o.ntt_fwd(i, a[r, i]) for i in range(S.K)]) for r in range(3)])
    assert np.array_equal(f, want)
    g = f.copy()
    ctx.ntt_inverse_host(g, 0, S.K)
    assert np.array_equal(g, a)
    # monomial X -> psi^(2 brev(k)+1) (SURVEY 9.2 sanity test)
    x = np.zeros((1, n), dtype=np.uint64)
    x[0, 1] = 1
    ctx.ntt_forward_host(x, 0, 1)
    assert int(x[0, 0]) == S.o.psi(0)


# ------------------------------------------------------------------ element-wise
@pytest.mark.parametrize("n,bits", [(8192, (60, 40, 40, 60)), (4096, (36, 36, 37))])
def test_elementwise_bit_exact(hg, n, bits):
    S = setup(n, bits)
    ctx = make_ctx(hg, S)
    L = S.Lmax
    rng = np.random.default_rng(11)
    B = 3
    a = rand_residues(rng, S.moduli[:L], (B, 2), n)
    b = rand_residues(rng, S.moduli[:L], (B, 2), n)
    c3 = rand_residues(rng, S.moduli[:L], (B, 3), n)
    pt = rand_residues(rng, S.moduli[:L], (B,), n)
    sc = 2.0**20
    A, Bc, C3 = ctx.upload_ct(a, sc), ctx.upload_ct(b, sc), ctx.upload_ct(c3, sc)
    P = ctx.upload_pt(pt, sc)
    out = ctx.ct(B)
    o = S.o

    def each(fn, *args):
        return np.stack([fn(*[x[i] for x in args]) for i in range(B)])

    ctx.add(out, A, Bc)
    assert np.array_equal(out.download(), each(o.add, a, b))
    ctx.sub(out, A, Bc)
    assert np.array_equal(out.download(), each(o.sub, a, b))
    ctx.negate(out, A)
    assert np.array_equal(out.download(), each(o.negate, a))
    # mixed sizes 2 vs 3
    ctx.add(out, A, C3)
    assert np.array_equal(out.download(), each(o.add, a, c3))
    ctx.sub(out, A, C3)
    assert np.array_equal(out.download(), each(o.sub, a, c3))
    ctx.sub(out, C3, A)
    assert np.array_equal(out.download(), each(o.sub, c3, a))
    ctx.add(out, C3, A)
    assert np.array_equal(out.download(), each(o.add, c3, a))
    # plaintext ops: per-element plaintexts (index -1) and one broadcast plaintext
    ctx.multiply_plain(out, A, P, -1)
    assert np.array_equal(out.download(), each(o.multiply_plain, a, pt))
    assert out.scale == sc * sc
    ctx.multiply_plain(out, C3, P, 1)
    assert np.array_equal(out.download(), np.stack([o.multiply_plain(c3[i], pt[1]) for i in range(B)]))
    ctx.add_plain(out, A, P, -1)
    assert np.array_equal(out.download(), each(o.add_plain, a, pt))
    ctx.sub_plain(out, A, P, 2)
    assert np.array_equal(out.download(), np.stack([o.sub_plain(a[i], pt[2]) for i in range(B)]))
    # tensor product and square
    ctx.multiply(out, A, Bc)
    assert np.array_equal(out.download(), each(o.multiply, a, b))
    ctx.square(out, A)
    assert np.array_equal(out.download(), each(o.square, a))
    # broadcast of a batch-1 second operand
    one = ctx.upload_ct(b[:1], sc)
    ctx.multiply(out, A, one)
    assert np.array_equal(out.download(), np.stack([o.multiply(a[i], b[0]) for i in range(B)]))
    # in place
    ctx.add(A, A, Bc)
    assert np.array_equal(A.download(), each(o.add, a, b))


def test_error_behaviour_matches_seal(hg):
    S = setup(8192, (60, 40, 40, 60))
    ctx = make_ctx(hg, S)
    rng = np.random.default_rng(3)
    a = rand_residues(rng, S.moduli[:3], (1, 2), S.n)
    A = ctx.upload_ct(a, 2.0**40)
    Bc = ctx.upload_ct(a, 2.0**41)
    out = ctx.ct(1)
    with pytest.raises(hg.InvalidArgument, match="scale mismatch"):
        ctx.add(out, A, Bc)
    low = ctx.upload_ct(a[:, :, :2], 2.0**40)
    with pytest.raises(hg.InvalidArgument, match="parameter mismatch"):
        ctx.add(out, A, low)
    big = ctx.upload_ct(a, 2.0**100)
    with pytest.raises(hg.InvalidArgument, match="scale out of bounds"):
        ctx.multiply(out, big, big)
    l1 = ctx.upload_ct(a[:, :, :1], 2.0**40)
    with pytest.raises(hg.InvalidArgument, match="end of modulus switching chain reached"):
        ctx.rescale_to_next(out, l1)
    with pytest.raises(hg.InvalidArgument, match="Galois key not present"):
        ctx.rotate_vector(out, A, 1)
    with pytest.raises(hg.InvalidArgument, match="step count too large"):
        ctx.rotate_vector(out, A, S.n // 2)
    ctx.multiply(out, A, A)
    with pytest.raises(hg.InvalidArgument, match="not enough relinearization keys"):
        ctx.relinearize(out, out)


def test_transparent_ciphertext_query(hg):
    """Ciphertext::is_transparent (SEAL 4.1): polynomials 1..size-1 all zero; the adapter turns a non-zero count
    into std::logic_error("result ciphertext is transparent") when the check is switched on."""
    S = setup(4096, (36, 36, 37))
    ctx = make_ctx(hg, S)
    rng = np.random.default_rng(4)
    a = rand_residues(rng, S.moduli[:2], (3, 2), S.n)
    a[1, 1] = 0  # second ciphertext: c1 = 0
    A = ctx.upload_ct(a, 2.0**20)
    assert ctx.transparent_count(A) == 1
    a[1, 1, 1, 77] = 5  # one non-zero word anywhere in c1 is enough
    assert ctx.transparent_count(ctx.upload_ct(a, 2.0**20)) == 0
    out = ctx.ct(3)
    ctx.sub(out, A, A)
    assert ctx.transparent_count(out) == 3
    a3 = rand_residues(rng, S.moduli[:2], (2, 3), S.n)
    a3[0, 2] = 0  # size 3 with c2 = 0 but c1 != 0: not transparent
    assert ctx.transparent_count(ctx.upload_ct(a3, 2.0**20)) == 0


# ------------------------------------------------------------------ rescale / key switching
@pytest.mark.parametrize("n,bits", [(8192, (60, 40, 40, 60)), (16384, (60, 31, 30, 30, 30, 60)), (4096, (36, 36, 37)),
                                    (32768, (60, 40, 40, 60)),
                                    # loader paths of the fused transforms: 50-bit digits on 30-bit limbs (direct conversion + one FP64
                                    # reduction), 30-bit digits on 50-bit limbs (integer policy without corrections), SEAL's 4_levels chain;
                                    # a 59 / 47 / 35-bit mix (every arithmetic policy in one chain); N = 32768 with small primes
                                    (8192, (50, 30, 30, 50, 50)), (16384, (59, 47, 35, 60)), (32768, (60, 30, 50, 60))])
def test_rescale_relin_galois_bit_exact(hg, n, bits):
    S = setup(n, bits)
    ctx = make_ctx(hg, S)
    rng = np.random.default_rng(5)
    o = S.o
    steps = [1, -1, 2, 8]
    gk = S.gk(steps)
    ctx.load_relin_key(S.rk)
    ctx.load_galois_keys(gk)
    B = 2
    for L in range(S.Lmax, 0, -1):
        mods = S.moduli[:L]
        c3 = rand_residues(rng, mods, (B, 3), n)
        c2 = np.ascontiguousarray(c3[:, :2])
        C3, C2 = ctx.upload_ct(c3, 2.0**30), ctx.upload_ct(c2, 2.0**30)
        out = ctx.ct(B)
        if L >= 2:
            ctx.rescale_to_next(out, C3)
            assert np.array_equal(out.download(), np.stack([o.rescale(c3[i]) for i in range(B)]))
            assert out.scale == 2.0**30 / mods[-1] and out.L == L - 1
            ctx.mod_switch_to_next(out, C3)
            assert np.array_equal(out.download(), np.stack([o.mod_switch(c3[i]) for i in range(B)]))
        ctx.relinearize(out, C3)
        assert np.array_equal(out.download(), np.stack([o.relinearize(c3[i], S.rk) for i in range(B)]))
        for st in steps:
            elt = orc.galois_elt_from_step(n, st)
            assert ctx.galois_elt_from_step(st) == elt
            ctx.rotate_vector(out, C2, st)
            assert np.array_equal(out.download(), np.stack([o.apply_galois(c2[i], elt, gk[elt]) for i in range(B)]))
        # in place + NAF chain: 7 = [-1, 8]
        want = np.stack([o.rotate(c2[i], 7, gk)[0] for i in range(B)])
        ctx.rotate_vector(C2, C2, 7)
        assert np.array_equal(C2.download(), want)
        # relinearize in place
        ctx.relinearize(C3, C3)
        assert np.array_equal(C3.download(), np.stack([o.relinearize(c3[i], S.rk) for i in range(B)]))


def test_rotation_zero_and_negative_naf(hg):
    S = setup(8192, (60, 40, 40, 60))
    ctx = make_ctx(hg, S)
    rng = np.random.default_rng(9)
    gk = S.gk([1, 2, 4, 8, -1, -2, -4, -8])
    ctx.load_galois_keys(gk)
    c2 = rand_residues(rng, S.moduli[:3], (1, 2), S.n)
    C2 = ctx.upload_ct(c2, 2.0**40)
    out = ctx.ct(1)
    ctx.rotate_vector(out, C2, 0)
    assert np.array_equal(out.download(), c2)
    for st in (3, -3, 5, -7):
        ctx.rotate_vector(out, C2, st)
        assert np.array_equal(out.download()[0], S.o.rotate(c2[0], st, gk)[0])


# ------------------------------------------------------------------ composites
def _diag_plaintexts(S, M, n1, n2, scale, L):
    dim = M.shape[0]
    slots = S.n // 2
    pts = np.empty((dim, L, S.n), dtype=np.uint64)
    for g in range(n2):
        for b in range(n1):
            d = g * n1 + b
            diag = np.array([M[r, (r + d) % dim] for r in range(dim)])
            pts[d] = S.enc.encode(np.roll(np.tile(diag, slots // dim), g * n1), scale, L)
    return pts


@pytest.mark.parametrize("n,dim,n1,n2", [(8192, 16, 4, 4), (16384, 32, 8, 4), (8192, 8, 8, 1), (8192, 8, 1, 8)])
def test_matvec_bsgs_bit_exact_and_decrypts(hg, n, dim, n1, n2):
    S = setup(n, (60, 40, 40, 60))
    ctx = make_ctx(hg, S)
    rng = np.random.default_rng(21)
    scale, L, B = 2.0**40, 3, 3
    M = rng.uniform(-1, 1, (dim, dim))
    V = rng.uniform(-1, 1, (B, dim))
    cts = np.stack([S.encrypt(np.tile(V[i], S.n // 2 // dim), scale, L, seed=i) for i in range(B)])
    pts = _diag_plaintexts(S, M, n1, n2, scale, L)
    bsteps = list(range(1, n1))
    gsteps = [g * n1 for g in range(1, n2)]
    gk = S.gk(bsteps + gsteps)
    ctx.load_galois_keys(gk)
    bk = [None] + [gk[orc.galois_elt_from_step(n, s)] for s in bsteps]
    gkeys = [None] + [gk[orc.galois_elt_from_step(n, s)] for s in gsteps]
    X = ctx.upload_ct(cts, scale)
    D = ctx.upload_pt(pts, scale)
    out = ctx.ct(B, 2)
    # exact chain of SEAL primitives, and every fast mode against its own oracle restatement
    for hoist, lazy in ((False, False), (True, True), (True, False), (False, True)):
        tol = ckks_tol(dim, n, scale, "hoisted" if hoist else "exact")
        want = S.o.matvec_bsgs(cts, n1, n2, pts, bk, gkeys, threads=4, hoist=hoist, lazy=lazy)
        ctx.matvec_bsgs(out, X, D, n1, n2, hoist=hoist, lazy=lazy)
        got = out.download()
        assert np.array_equal(got, want)
        for i in range(B):
            dec = S.decrypt(got[i], out.scale).real[:dim]
            assert np.max(np.abs(dec - M @ V[i])) < tol
    # without the final rescale (multi-GPU partial sums): same ciphertext before rescale
    ctx.matvec_bsgs(out, X, D, n1, n2, rescale=False, hoist=True, lazy=True)
    assert out.L == L
    part = out.download()
    want = S.o.matvec_bsgs(cts, n1, n2, pts, bk, gkeys, threads=4, fast=True)
    assert np.array_equal(np.stack([S.o.rescale(part[i]) for i in range(B)]), want)


@pytest.mark.parametrize("n1,n2,L", [(2, 2, 2), (2, 4, 3), (4, 4, 3)])
def test_matvec_bsgs_range_all_rotated_giant_steps(hg, n1, n2, L):
    """A diagonal shard with g_first > 0 rotates EVERY giant step of the call (nrot = n2), one more key-switch
    group than a g_first = 0 call of the same shape: the scratch plan used to be one group short when n2 >= n1
    (silent arena overrun).  Fresh context per case (nothing grown by an earlier call), every non-DH
    mode against the oracle's restatement, no final rescale (multi-GPU partial sums) and with it."""
    n, B = 8192, 3
    S = setup(n, (60, 40, 40, 60))
    rng = np.random.default_rng(27)
    scale = 2.0**40
    cts = rand_residues(rng, S.moduli[:L], (B, 2), n)
    g0 = n2  # the second shard of a 2 * n2 giant-step matvec
    pts = rand_residues(rng, S.moduli[:L], (n1 * n2,), n)
    bsteps, gsteps = list(range(1, n1)), [(g0 + g) * n1 for g in range(n2)]
    gk = S.gk(bsteps + gsteps)
    bk = [None] + [gk[orc.galois_elt_from_step(n, s)] for s in bsteps]
    gkeys = [gk[orc.galois_elt_from_step(n, s)] for s in gsteps]
    for hoist, lazy, rescale in ((True, True, False), (False, False, False), (False, True, True), (True, False, True)):
        ctx = hg.Context(n, S.moduli)  # fresh arena: the first call must size it correctly on its own
        ctx.load_galois_keys(gk)
        X = ctx.upload_ct(cts, scale, size_cap=2, L_cap=L)
        D = ctx.upload_pt(pts, scale)
        out = ctx.ct(B, 2, L)
        want = S.o.matvec_bsgs(cts, n1, n2, pts, bk, gkeys, threads=4, hoist=hoist, lazy=lazy, rescale=rescale, g_first=g0)
        ctx.matvec_bsgs(out, X, D, n1, n2, hoist=hoist, lazy=lazy, rescale=rescale, g_first=g0)
        assert np.array_equal(out.download(), want), (hoist, lazy, rescale)


def _diag_plaintexts_ext(S, M, n1, n2, scale, L):
    dim = M.shape[0]
    slots = S.n // 2
    pts = np.empty((dim, L + 1, S.n), dtype=np.uint64)
    for g in range(n2):
        for b in range(n1):
            d = g * n1 + b
            diag = np.array([M[r, (r + d) % dim] for r in range(dim)])
            pts[d] = S.enc.encode_ext(np.roll(np.tile(diag, slots // dim), g * n1), scale, L)
    return pts


@pytest.mark.parametrize("n,dim,n1,n2", [(8192, 16, 4, 4), (16384, 32, 8, 4), (8192, 32, 32, 1), (8192, 8, 1, 8), (8192, 48, 24, 2),
                                         (8192, 64, 32, 2), (32768, 16, 4, 4), (4096, 8, 4, 2)])
def test_matvec_bsgs_double_hoisted(hg, n, dim, n1, n2):
    """HEGPU_MATVEC_DH against its oracle restatement (orc_matvec_bsgs_dh, itself pinned by a big-integer
    restatement in tests/test_oracle_evaluator.py): bit-exact, decrypts to M @ v, with and without the
    final rescale, and diagonal-sharded (g_first > 0)."""
    S = setup(n, (60, 40, 40, 60))
    ctx = make_ctx(hg, S)
    rng = np.random.default_rng(22)
    scale, L, B = 2.0**40, 3, 3
    M = rng.uniform(-1, 1, (dim, dim))
    V = rng.uniform(-1, 1, (B, dim))
    tiles = (S.n // 2) % dim == 0  # otherwise (n1 = 24): random plaintexts, bit-exactness only
    cts = np.stack([S.encrypt(np.tile(V[i], S.n // 2 // dim) if tiles else V[i], scale, L, seed=i) for i in range(B)])
    if tiles:
        ptsx = _diag_plaintexts_ext(S, M, n1, n2, scale, L)
    else:
        ptsx = rand_residues(rng, S.moduli[:L] + [S.moduli[-1]], (n1 * n2,), n)
    bsteps = list(range(1, n1))
    gsteps = [g * n1 for g in range(1, n2)]
    gk = S.gk(bsteps + gsteps)
    ctx.load_galois_keys(gk)
    bk = [None] + [gk[orc.galois_elt_from_step(n, s)] for s in bsteps]
    gkeys = [None] + [gk[orc.galois_elt_from_step(n, s)] for s in gsteps]
    X = ctx.upload_ct(cts, scale)
    D = ctx.upload_pt_ext(ptsx, scale)
    out = ctx.ct(B, 2)
    tol = ckks_tol(dim, n, scale, "hoisted")
    want = S.o.matvec_bsgs(cts, n1, n2, ptsx, bk, gkeys, threads=4, dh=True)
    ctx.matvec_bsgs(out, X, D, n1, n2, dh=True)
    got = out.download()
    assert np.array_equal(got, want)
    for i in range(B if tiles else 0):
        dec = S.decrypt(got[i], out.scale).real[:dim]
        assert np.max(np.abs(dec - M @ V[i])) < tol
    ctx.matvec_bsgs(out, X, D, n1, n2, rescale=False, dh=True)
    assert out.L == L
    assert np.array_equal(out.download(), S.o.matvec_bsgs(cts, n1, n2, ptsx, bk, gkeys, threads=4, dh=True, rescale=False))
    if n2 >= 2:  # the giant steps of the second half as their own call (diagonal sharding)
        g0 = n2 // 2
        cnt = n2 - g0
        Dp = ctx.upload_pt_ext(np.ascontiguousarray(ptsx[g0 * n1:]), scale)
        ctx.matvec_bsgs(out, X, Dp, n1, cnt, rescale=False, dh=True, g_first=g0)
        wantp = S.o.matvec_bsgs(cts, n1, cnt, ptsx[g0 * n1:], bk, gkeys[g0:], threads=4, dh=True, rescale=False, g_first=g0)
        assert np.array_equal(out.download(), wantp)
    # the ordinary plaintext set is rejected in DH mode and vice versa (no silent misuse)
    Dn = ctx.upload_pt(np.ascontiguousarray(ptsx[:, :L]), scale)
    with pytest.raises(hg.InvalidArgument):
        ctx.matvec_bsgs(out, X, Dn, n1, n2, dh=True)
    with pytest.raises(hg.InvalidArgument):
        ctx.matvec_bsgs(out, X, D, min(n1, 16), n2, hoist=True)


@pytest.mark.parametrize("bits,L", [((60, 40, 40, 60), 2), ((60, 40, 40, 40, 60), 4), ((60, 40, 40, 40, 60), 3), ((50, 50, 60), 2), ((40, 60), 1),
                                    ((50, 40, 40, 50, 50), 4), ((59, 47, 41, 60), 3)])
def test_matvec_double_hoisted_levels(hg, bits, L):
    """Double-hoisted matvec below the top level (the special prime is then NOT the limb after the
    ciphertext's last one), with four digits, with 50-bit primes only (integer policy everywhere) and with
    a single digit: bit-exact against the oracle, decrypts to M @ v."""
    n, dim, n1, n2, B = 8192, 16, 4, 4, 2
    S = setup(n, bits)
    ctx = make_ctx(hg, S)
    rng = np.random.default_rng(24)
    scale = 2.0**40 if L > 1 else 2.0**15
    M = rng.uniform(-1, 1, (dim, dim))
    V = rng.uniform(-1, 1, (B, dim))
    cts = np.stack([S.encrypt(np.tile(V[i], n // 2 // dim), scale, L, seed=i) for i in range(B)])
    ptsx = _diag_plaintexts_ext(S, M, n1, n2, scale, L)
    bsteps, gsteps = list(range(1, n1)), [g * n1 for g in range(1, n2)]
    gk = S.gk(bsteps + gsteps)
    ctx.load_galois_keys(gk)
    bk = [None] + [gk[orc.galois_elt_from_step(n, s)] for s in bsteps]
    gkeys = [None] + [gk[orc.galois_elt_from_step(n, s)] for s in gsteps]
    X = ctx.upload_ct(cts, scale)
    D = ctx.upload_pt_ext(ptsx, scale)
    out = ctx.ct(B, 2)
    rescale = L > 1
    want = S.o.matvec_bsgs(cts, n1, n2, ptsx, bk, gkeys, threads=4, dh=True, rescale=rescale)
    ctx.matvec_bsgs(out, X, D, n1, n2, dh=True, rescale=rescale)
    got = out.download()
    assert out.L == (L - 1 if rescale else L)
    assert np.array_equal(got, want)
    if rescale:
        tol = ckks_tol(dim, n, scale, "hoisted", out_scale=out.scale)
        for i in range(B):
            assert np.max(np.abs(S.decrypt(got[i], out.scale).real[:dim] - M @ V[i])) < tol


@pytest.mark.parametrize("n,bits,n1,n2,B,g_first", [(8192, (60, 40, 40, 60), 8, 2, 5, 0), (16384, (60, 40, 40, 60), 32, 4, 21, 0),
                                                    (8192, (60, 40, 40, 60), 4, 6, 18, 0), (8192, (50, 50, 50, 60), 16, 3, 3, 2)])
def test_matvec_dh_integer_mma_mode_is_bit_identical(hg, n, bits, n1, n2, B, g_first):
    """HEGPU_MATVEC_IMMA (opt-in, experimental): the fused inner sums as 8-bit-limb integer matrix products on the
    warp-level MMA units must return exactly the bits of the default (IMAD / FP64) kernel -- which the tests above pin to
    the oracle -- for full and ragged ciphertext tiles, one and several k-steps, several launch groups of giant steps,
    a sharded giant-step range, 40- and 50/60-bit limbs; a second call reuses the pre-multiplied diagonals."""
    S = setup(n, bits)
    ctx = make_ctx(hg, S)
    rng = np.random.default_rng(77)
    L, scale = 3, 2.0**40
    cts = rand_residues(rng, S.moduli[:L], (B, 2), n)
    ptsx = rand_residues(rng, S.moduli[:L] + [S.moduli[-1]], (n1 * n2,), n)
    steps = list(range(1, n1)) + [g * n1 for g in range(max(g_first, 1), g_first + n2)]
    ctx.load_galois_keys(S.gk(steps))
    X, D = ctx.upload_ct(cts, scale), ctx.upload_pt_ext(ptsx, scale)
    ref_out, out = ctx.ct(B, 2), ctx.ct(B, 2)
    ctx.matvec_bsgs(ref_out, X, D, n1, n2, dh=True, g_first=g_first)
    want = ref_out.download()
    for _ in range(2):
        ctx.matvec_bsgs(out, X, D, n1, n2, dh=True, g_first=g_first, imma=True)
        assert np.array_equal(out.download(), want)
    # new diagonals invalidate the cached products
    ptsx2 = rand_residues(rng, S.moduli[:L] + [S.moduli[-1]], (n1 * n2,), n)
    D.upload_ext(ptsx2, scale)
    ctx.matvec_bsgs(ref_out, X, D, n1, n2, dh=True, g_first=g_first)
    ctx.matvec_bsgs(out, X, D, n1, n2, dh=True, g_first=g_first, imma=True)
    assert np.array_equal(out.download(), ref_out.download())


def test_matvec_double_hoisted_two_execution_slots(hg):
    """Batches of 32 or more ciphertexts are split into chunks that run on two streams with separate
    scratch (and an uneven last chunk): every ciphertext must still match the oracle bit for bit, also
    when the same context is used again right away (scratch reuse across calls)."""
    n, n1, n2, L, B = 4096, 4, 3, 2, 41
    S = setup(n, (36, 36, 37))
    ctx = make_ctx(hg, S)
    rng = np.random.default_rng(23)
    scale = 2.0**15
    cts = rand_residues(rng, S.moduli[:L], (B, 2), n)
    ptsx = rand_residues(rng, S.moduli[:L] + [S.moduli[-1]], (n1 * n2,), n)
    bsteps, gsteps = list(range(1, n1)), [g * n1 for g in range(1, n2)]
    gk = S.gk(bsteps + gsteps)
    ctx.load_galois_keys(gk)
    bk = [None] + [gk[orc.galois_elt_from_step(n, s)] for s in bsteps]
    gkeys = [None] + [gk[orc.galois_elt_from_step(n, s)] for s in gsteps]
    X = ctx.upload_ct(cts, scale, size_cap=2, L_cap=L)
    D = ctx.upload_pt_ext(ptsx, scale)
    out, out2 = ctx.ct(B, 2, L), ctx.ct(B, 2, L)
    want = S.o.matvec_bsgs(cts, n1, n2, ptsx, bk, gkeys, threads=4, dh=True)
    ctx.matvec_bsgs(out, X, D, n1, n2, dh=True)
    ctx.matvec_bsgs(out2, X, D, n1, n2, dh=True)  # enqueued while the first call may still be running
    assert np.array_equal(out.download(), want)
    assert np.array_equal(out2.download(), want)


@pytest.mark.parametrize("case_b", [False, True])
def test_bmatmul_reference_loop_order(hg, case_b):
    """BatchedMatrix::matmul (he_linalg.cpp:943-1006) restated on the oracle in the reference's own
    loop order with SEAL's default power-of-two Galois keys (NAF rotations)."""
    n, dim = 8192, 4
    S = setup(n, (60, 40, 40, 60))
    ctx = make_ctx(hg, S)
    rng = np.random.default_rng(31)
    scale, L = 2.0**40, 3
    gk = S.gk([1, 2, 4, -1, -2, -4])
    ctx.load_galois_keys(gk)
    ctx.load_relin_key(S.rk)
    A = rng.uniform(-1, 1, (dim, dim))
    Bm = rng.uniform(-1, 1, (dim, dim))
    slots = n // 2
    rep = lambda v: np.tile(v, slots // dim)
    if not case_b:
        this_vals = [np.array([A[r, (r + d) % dim] for r in range(dim)]) for d in range(dim)]  # diagonals of A
        other_vals = [Bm[:, j] for j in range(dim)]  # columns of B
    else:
        this_vals = [A[:, j] for j in range(dim)]  # columns of A
        other_vals = [Bm[j, :] for j in range(dim)]  # rows of B (transposed col batching)
    this_c = np.stack([S.encrypt(rep(v), scale, L, seed=10 + i) for i, v in enumerate(this_vals)])
    other_c = np.stack([S.encrypt(rep(v), scale, L, seed=20 + i) for i, v in enumerate(other_vals)])
    o = S.o
    want = []
    for i in range(dim):
        acc = None
        for j in range(dim):
            src, st = (other_c[i], j) if not case_b else (other_c[j], i)
            r, _ = o.rotate(src, st, gk)
            t = o.multiply(r, this_c[j])
            acc = t if acc is None else o.add(acc, t)
        want.append(o.rescale(o.relinearize(acc, S.rk)))
    want = np.stack(want)
    T, O = ctx.upload_ct(this_c, scale), ctx.upload_ct(other_c, scale)
    out = ctx.ct(dim, 2)
    ctx.bmatmul(out, T, O, dim, dim, case_b)
    got = out.download()
    assert np.array_equal(got, want)
    tol = ckks_tol(dim, n, scale)
    if not case_b:
        for j in range(dim):  # column j of A @ B
            assert np.max(np.abs(S.decrypt(got[j], out.scale).real[:dim] - (A @ Bm)[:, j])) < tol


def test_matmul_elemwise_bit_exact(hg):
    """Matrix::matmul (he_linalg.cpp:202-236): one ciphertext per entry, column-major."""
    n = 4096
    S = setup(n, (36, 36, 37))
    ctx = make_ctx(hg, S)
    ctx.load_relin_key(S.rk)
    rng = np.random.default_rng(41)
    rows, inner, cols, L = 2, 3, 2, 2
    a = rand_residues(rng, S.moduli[:L], (rows * inner, 2), n)
    b = rand_residues(rng, S.moduli[:L], (inner * cols, 2), n)
    o = S.o
    want = []
    for j in range(cols):
        for i in range(rows):
            acc = None
            for k in range(inner):
                t = o.multiply(a[i + k * rows], b[k + j * inner])
                acc = t if acc is None else o.add(acc, t)
            want.append(o.rescale(o.relinearize(acc, S.rk)))
    want = np.stack(want)  # column-major index i + j*rows
    A, Bc = ctx.upload_ct(a, 2.0**15), ctx.upload_ct(b, 2.0**15)
    out = ctx.ct(rows * cols, 2)
    ctx.matmul_elemwise(out, A, Bc, rows, inner, cols)
    assert np.array_equal(out.download(), want)


def test_bfft_stage_and_fft_butterflies_bit_exact(hg):
    n = 8192
    S = setup(n, (60, 40, 40, 60))
    ctx = make_ctx(hg, S)
    gk = S.gk([4, -4])
    ctx.load_galois_keys(gk)
    rng = np.random.default_rng(51)
    L, sc = 3, 2.0**40
    o = S.o
    y = rand_residues(rng, S.moduli[:L], (2, 2), n)
    d = rand_residues(rng, S.moduli[:L], (3,), n)
    for with_d2 in (True, False):
        Y = ctx.upload_ct(y, sc)
        ctx.bfft_stage(Y, ctx.upload_pt(d, sc), 4, with_d2)
        want = []
        for i in range(2):
            y0 = o.rescale(o.multiply_plain(y[i], d[0]))
            y1 = o.rescale(o.multiply_plain(o.rotate(y[i], 4, gk)[0], d[1]))
            r = o.add(y0, y1)
            if with_d2:
                r = o.add(r, o.rescale(o.multiply_plain(o.rotate(y[i], -4, gk)[0], d[2])))
            want.append(r)
        assert np.array_equal(Y.download(), np.stack(want))
    half = 3
    ev = rand_residues(rng, S.moduli[:L], (half, 2), n)
    od = rand_residues(rng, S.moduli[:L], (half, 2), n)
    w = rand_residues(rng, S.moduli[:L], (half,), n)
    one = rand_residues(rng, S.moduli[:L], (1,), n)
    out = ctx.ct(2 * half, 2)
    ctx.fft_butterflies(out, ctx.upload_ct(ev, sc), ctx.upload_ct(od, sc), ctx.upload_pt(w, sc), ctx.upload_pt(one, sc))
    got = out.download()
    for k in range(half):
        t = o.rescale(o.multiply_plain(od[k], w[k]))
        e = o.rescale(o.multiply_plain(ev[k], one[0]))
        assert np.array_equal(got[k], o.add(e, t))
        assert np.array_equal(got[k + half], o.sub(e, t))


def test_reduce_fixup(hg):
    S = setup(8192, (60, 40, 40, 60))
    ctx = make_ctx(hg, S)
    rng = np.random.default_rng(61)
    parts = rand_residues(rng, S.moduli[:3], (8, 1, 2), S.n)
    summed = parts.sum(axis=0, dtype=np.uint64)  # what an NCCL uint64 sum produces
    T = ctx.upload_ct(summed, 1.0)
    ctx.reduce_fixup(T, 8)
    want = parts[0]
    for k in range(1, 8):
        want = np.stack([S.o.add(want[0], parts[k][0])])
    assert np.array_equal(T.download(), want)


def test_async_copies_pipeline(hg):
    """hegpu_ct_upload_async / download_async: double-buffered pipeline equals the synchronous path."""
    import torch

    S = setup(8192, (60, 40, 40, 60))
    ctx = make_ctx(hg, S)
    rng = np.random.default_rng(71)
    L, B, sc = 3, 4, 2.0**40
    ins = [torch.from_numpy(rand_residues(rng, S.moduli[:L], (B, 2), S.n)).pin_memory() for _ in range(4)]
    outs = [torch.empty((B, 2, L, S.n), dtype=torch.int64).pin_memory() for _ in range(2)]
    X = [ctx.ct(B, 2, L) for _ in range(2)]
    O = [ctx.ct(B, 2, L) for _ in range(2)]
    results = []
    X[0].upload_async(ins[0].data_ptr(), sc, 2, L)
    for i in range(4):
        cur, nxt = i & 1, (i + 1) & 1
        if i + 1 < 4:
            X[nxt].upload_async(ins[i + 1].data_ptr(), sc, 2, L)
        if i >= 2:
            O[cur].copy_wait()
            results.append(outs[cur].numpy().view(np.uint64).copy())
        ctx.add(O[cur], X[cur], X[cur])
        O[cur].download_async(outs[cur].data_ptr())
    for k in (0, 1):
        O[k].copy_wait()
        results.append(outs[k].numpy().view(np.uint64).copy())
    ctx.sync()
    for i in range(4):
        a = ins[i].numpy()
        want = np.stack([S.o.add(a[b], a[b]) for b in range(B)])
        assert np.array_equal(results[i], want), i
