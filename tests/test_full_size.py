"""BASELINE.json's configurations at FULL size, through size-independent properties (decrypt and
compare with float64 numpy within the CKKS tolerance) and, where the oracle finishes in seconds,
bit for bit:

  cfg 3  N = 16384, encrypted x encrypted 64x64 matrix product (BatchedMatrix::matmul case B,
         he_linalg.cpp:943-1006) with relinearize + rescale, through the C++ host mirror;
  cfg 4  homomorphic FFT over 4096 slots at N = 16384 (he::fft::bfft, he_fft.cpp:166-223), 12 stages,
         23 rotations, 13-level chain {60, 40 x 12, 60}, through the C++ host mirror, replayed on the
         oracle with the plaintext limbs the host encoded.  (The reference encodes every stage diagonal
         at the ciphertext's current scale, so the scale follows s <- s^2 / q_last: the log-deviation of
         s from the primes DOUBLES per stage.  Twelve stages therefore need primes within ~1e-5 of each
         other, which 26-bit primes at N = 16384 -- the only way to fit 12 levels under SEAL's 438-bit
         128-bit-security bound -- cannot give: that chain ends in "scale out of bounds" on SEAL and
         here alike.  40-bit primes work; the 600-bit chain needs sec_level_type::none in SEAL.)
  cfg 5  N = 32768, 512x512 plaintext-diagonal matvec, sharded by diagonals over two ranks
         (emulated on one GPU), double-hoisted 32 x 16, against the oracle's sharded restatement.

  cfg 1  N = 8192 {60,40,40,60}, 16x16 encrypted-diagonal matrix x encrypted vector (case A), SEAL's default
         power-of-two Galois keys -> NAF chains, the reference's loop order (matrix_operations.cpp:1042-1175);
  cfg 2  N = 16384, 128x128 plaintext diagonals x encrypted vector: the double-hoisted 32 x 4 mode the bench
         times AND the exact mode (chain of SEAL primitives, 16 x 8), each against the oracle bit for bit;
  cfg 3' Matrix::matmul 64x64 (he_linalg.cpp:202-236), 2 x 4096 ciphertexts;
  the shipped demos fft.cpp:127-241 (bfft, n = 128, {60,31,30x9,60}, scale 2^30, inputs i + 7.1) and
  fft.cpp:13-125 (fft over 128 ciphertexts, {60,30x10,60}); Matrix::matmul_pow (he_linalg.cpp:316-349)."""
import numpy as np
import pytest
import torch

import hegpu_loader
from fixtures import ckks_tol, setup
from oracle import oracle as orc
from test_host_cpp import run

pytestmark = pytest.mark.gpu


def test_cfg3_encrypted_matmul_64x64(tmp_path):
    n, dim = 16384, 64
    S = setup(n, (60, 40, 40, 60))
    rng = np.random.default_rng(33)
    L, sc = 3, 2.0**40
    A = rng.uniform(-1, 1, (dim, dim))
    Bm = rng.uniform(-1, 1, (dim, dim))
    reps = n // 2 // dim
    this_c = [S.encrypt(np.tile(A[:, j], reps), sc, L, seed=j) for j in range(dim)]        # columns of A
    other_c = [S.encrypt(np.tile(Bm[j, :], reps), sc, L, seed=100 + j) for j in range(dim)]  # transposed columns of B
    steps = [s * (1 << k) for k in range(7) for s in (1, -1)]
    outs, _, gk, _ = run(tmp_path, S, "bmatmul", [1, dim, dim, dim], cts=[(x, sc) for x in this_c + other_c], rk=True, gk_steps=steps)
    assert len(outs) == dim
    C = A @ Bm
    tol = ckks_tol(3 * dim, n, sc)  # <= 3 NAF key-switches per product, 64 products per output
    r = np.arange(dim)
    worst = 0.0
    for i, (ct, scale) in enumerate(outs):
        assert ct.shape == (2, L - 1, n)
        dec = S.decrypt(ct, scale).real[:dim]
        worst = max(worst, float(np.max(np.abs(dec - C[r, (r + i) % dim]))))  # result is diagonal-batched
    assert worst < tol
    # spot check one output bit for bit against the oracle driven in the reference's loop order
    i = 5
    acc = None
    for j in range(dim):
        t = S.o.multiply(S.o.rotate(other_c[j], i, gk)[0], this_c[j])
        acc = t if acc is None else S.o.add(acc, t)
    assert np.array_equal(outs[i][0], S.o.rescale(S.o.relinearize(acc, S.rk)))


def test_cfg4_bfft_4096_slots(tmp_path):
    n, m = 16384, 4096
    S = setup(n, (60,) + (40,) * 12 + (60,))
    sc, L = 2.0**40, 13
    rng = np.random.default_rng(44)
    data = rng.uniform(-1, 1, m) / 64.0
    ct = S.encrypt(np.tile(data, n // 2 // m), sc, L, seed=3)
    steps = [s * (1 << k) for k in range(12) for s in (1, -1)]
    outs, pts, gk, _ = run(tmp_path, S, "bfft", [m, 0], cts=[(ct, sc)], gk_steps=steps)
    o = S.o
    y, pi = ct, 0
    for i in range(1, 13):
        st, with_d2 = m >> i, i != 1
        y0 = o.rescale(o.multiply_plain(y, pts[pi][0]))
        y1 = o.rescale(o.multiply_plain(o.rotate(y, st, gk)[0], pts[pi + 1][0]))
        nxt = o.add(y0, y1)
        if with_d2:
            nxt = o.add(nxt, o.rescale(o.multiply_plain(o.rotate(y, -st, gk)[0], pts[pi + 2][0])))
        pi += 3 if with_d2 else 2
        y = nxt
    got, got_scale = outs[0]
    assert got.shape == (2, 1, n)          # 12 levels consumed
    assert np.array_equal(got, y)          # bit-exact replay with the host's plaintext limbs
    dec = S.decrypt(got, got_scale)[:m]
    ref = np.fft.fft(data)
    brev = np.array([int(format(i, "012b")[::-1], 2) for i in range(m)])
    err = np.max(np.abs(dec - ref[brev]))  # output is in bit-reversed order (fft.cpp:224-238)
    assert err < 1e-3 * max(1.0, float(np.max(np.abs(ref)))), err


def test_cfg5_diag_sharded_512x512_n32768():
    hg = hegpu_loader.load()
    from hegpu_b200.multigpu import as_torch_i64, giant_step_range

    n, dim, n1, n2, B, world = 32768, 512, 32, 16, 2, 2
    S = setup(n, (60, 40, 40, 60))
    L, sc = 3, 2.0**40
    rng = np.random.default_rng(55)
    M = rng.uniform(-1, 1, (dim, dim))
    V = rng.uniform(-1, 1, (B, dim))
    slots = n // 2
    cts = np.stack([S.encrypt(np.tile(V[i], slots // dim), sc, L, seed=i) for i in range(B)])
    r = np.arange(dim)
    ptsx = np.empty((dim, L + 1, n), dtype=np.uint64)
    for g in range(n2):
        for k in range(n1):
            d = g * n1 + k
            ptsx[d] = S.enc.encode_ext(np.roll(np.tile(M[r, (r + d) % dim], slots // dim), g * n1), sc, L)
    bsteps, gsteps = list(range(1, n1)), [g * n1 for g in range(1, n2)]
    gk = S.gk(bsteps + gsteps)
    bk = [None] + [gk[orc.galois_elt_from_step(n, s)] for s in bsteps]
    gkeys = [None] + [gk[orc.galois_elt_from_step(n, s)] for s in gsteps]
    ctx = hg.Context(n, S.moduli)
    ctx.load_galois_keys(gk)
    X = ctx.upload_ct(cts, sc, size_cap=2, L_cap=L)
    parts, want_sum = [], None
    for rank in range(world):
        g0, cnt = giant_step_range(n2, world, rank)
        sl = np.ascontiguousarray(ptsx[g0 * n1:(g0 + cnt) * n1])
        D = ctx.upload_pt_ext(sl, sc)
        p = ctx.ct(B, 2, L)
        ctx.matvec_bsgs(p, X, D, n1, cnt, rescale=False, dh=True, g_first=g0)
        want = S.o.matvec_bsgs(cts, n1, cnt, sl, bk, gkeys[g0:g0 + cnt], threads=4, dh=True, rescale=False, g_first=g0)
        assert np.array_equal(p.download(), want)
        want_sum = want if want_sum is None else np.stack([S.o.add(want_sum[b], want[b]) for b in range(B)])
        parts.append(p)
    ctx.sync()
    acc = as_torch_i64(parts[0])
    acc += as_torch_i64(parts[1])  # stands in for the NCCL uint64 all-reduce
    torch.cuda.synchronize()
    ctx.reduce_fixup(parts[0], world)
    out = ctx.ct(B, 2, L - 1)
    ctx.rescale_to_next(out, parts[0])
    got = out.download()
    assert np.array_equal(got, np.stack([S.o.rescale(want_sum[b]) for b in range(B)]))
    tol = ckks_tol(dim, n, sc, "hoisted")
    for b in range(B):
        dec = S.decrypt(got[b], out.scale).real[:dim]
        assert np.max(np.abs(dec - M @ V[b])) < tol


def _diag_setup(S, n, dim, n1, n2, B, L, sc, seed):
    rng = np.random.default_rng(seed)
    M = rng.uniform(-1, 1, (dim, dim))
    V = rng.uniform(-1, 1, (B, dim))
    slots = n // 2
    cts = np.stack([S.encrypt(np.tile(V[i], slots // dim), sc, L, seed=i) for i in range(B)])
    r = np.arange(dim)
    ptsx = np.empty((dim, L + 1, n), dtype=np.uint64)
    for g in range(n2):
        for k in range(n1):
            d = g * n1 + k
            ptsx[d] = S.enc.encode_ext(np.roll(np.tile(M[r, (r + d) % dim], slots // dim), g * n1), sc, L)
    bsteps, gsteps = list(range(1, n1)), [g * n1 for g in range(1, n2)]
    gk = S.gk(bsteps + gsteps)
    bk = [None] + [gk[orc.galois_elt_from_step(n, st)] for st in bsteps]
    gkeys = [None] + [gk[orc.galois_elt_from_step(n, st)] for st in gsteps]
    return M, V, cts, ptsx, gk, bk, gkeys


@pytest.mark.parametrize("mode,n1,n2", [("dh", 32, 4), ("exact", 16, 8), ("hoist_lazy", 16, 8)])
def test_cfg2_matvec_128x128_n16384(mode, n1, n2):
    """BASELINE cfg 2 at full size: N = 16384 {60,40,40,60}, 128 x 128 plaintext
    diagonals, batch 6, through the C ABI.  `exact` is the chain of SEAL primitives (the only mode that is a
    sequence of seal::Evaluator calls: rotate_vector, multiply_plain, add, rescale); `dh` is what bench.py
    times.  Bit-exact against the oracle's restatement of the same mode, every ciphertext decrypted."""
    hg = hegpu_loader.load()
    n, dim, B, L, sc = 16384, 128, 6, 3, 2.0**40
    S = setup(n, (60, 40, 40, 60))
    M, V, cts, ptsx, gk, bk, gkeys = _diag_setup(S, n, dim, n1, n2, B, L, sc, seed=202)
    ctx = hg.Context(n, S.moduli)
    ctx.load_galois_keys(gk)
    X = ctx.upload_ct(cts, sc, size_cap=2, L_cap=L)
    out = ctx.ct(B, 2, L)
    if mode == "dh":
        D = ctx.upload_pt_ext(ptsx, sc)
        ctx.matvec_bsgs(out, X, D, n1, n2, dh=True)
        want = S.o.matvec_bsgs(cts, n1, n2, ptsx, bk, gkeys, threads=8, dh=True)
    else:
        pts = np.ascontiguousarray(ptsx[:, :L])
        D = ctx.upload_pt(pts, sc)
        fast = mode == "hoist_lazy"
        ctx.matvec_bsgs(out, X, D, n1, n2, hoist=fast, lazy=fast)
        want = S.o.matvec_bsgs(cts, n1, n2, pts, bk, gkeys, threads=8, fast=fast)
    got = out.download()
    assert out.L == L - 1 and got.shape == (B, 2, L - 1, n)
    assert np.array_equal(got, want)
    tol = ckks_tol(dim, n, sc, "exact" if mode == "exact" else "hoisted")
    worst = max(float(np.max(np.abs(S.decrypt(got[i], out.scale).real[:dim] - M @ V[i]))) for i in range(B))
    print(f"cfg2 {mode}: max |err| = {worst:.3e} (tol {tol:.3e})")
    assert worst < tol


def test_cfg1_16x16_case_a_n8192(tmp_path):
    """BASELINE cfg 1 = the reference's own demo shape (matrix_operations.cpp:1042-1175 parameters: N = 8192,
    {60,40,40,60}, scale 2^40) as a 16x16 matrix x vector, case A of BatchedMatrix::matmul (he_linalg.cpp:943-1006):
    16 ENCRYPTED generalized diagonals, one encrypted vector, Galois keys for the powers of two only (what
    create_galois_keys() without a step list provides), so steps 3, 5, 6, 7, ... run as SEAL's NAF chains.
    Through the C++ host mirror; bit-exact against the oracle driven in the reference's loop order."""
    n, dim = 8192, 16
    S = setup(n, (60, 40, 40, 60))
    rng = np.random.default_rng(101)
    L, sc = 3, 2.0**40
    M = rng.uniform(-1, 1, (dim, dim))
    v = rng.uniform(-1, 1, dim)
    reps = n // 2 // dim
    r = np.arange(dim)
    diags = [S.encrypt(np.tile(M[r, (r + j) % dim], reps), sc, L, seed=j) for j in range(dim)]
    vec = S.encrypt(np.tile(v, reps), sc, L, seed=99)
    steps = [sg * (1 << k) for k in range(5) for sg in (1, -1)]  # the default key set restricted to |step| <= 16
    outs, _, gk, _ = run(tmp_path, S, "bmatmul", [0, dim, 1, dim], cts=[(x, sc) for x in diags] + [(vec, sc)], rk=True, gk_steps=steps)
    assert len(outs) == 1
    o = S.o
    acc, nks = None, 0
    for j in range(dim):
        rot, k = o.rotate(vec, j, gk)
        nks += k
        t = o.multiply(rot, diags[j])
        acc = t if acc is None else o.add(acc, t)
    assert nks == 28  # 28 key-switches for the 15 non-trivial rotations of a 16-diagonal matvec
    got, got_scale = outs[0]
    assert np.array_equal(got, o.rescale(o.relinearize(acc, S.rk)))
    err = float(np.max(np.abs(S.decrypt(got, got_scale).real[:dim] - M @ v)))
    print(f"cfg1: max |err| = {err:.3e}")
    assert err < ckks_tol(2 * dim, n, sc)  # up to two NAF key-switches per term, plus the relinearisation


def test_cfg3_matrix_matmul_64x64():
    """BASELINE cfg 3 (ii): Matrix::matmul (he_linalg.cpp:202-236), one ciphertext per matrix entry, 64 x 64 x 64 at
    N = 16384: 2 x 4096 input ciphertexts (3 GiB each), 262 144 ct x ct products, 4 096 relinearisations and
    rescales, through the C ABI composite.  Sampled outputs bit for bit against the oracle in the reference's order
    (C(i,j) = relin_rescale(sum_k A(i,k) B(k,j)), column-major), every sampled output decrypted against numpy."""
    hg = hegpu_loader.load()
    n, d, L, sc = 16384, 64, 3, 2.0**40
    S = setup(n, (60, 40, 40, 60))
    rng = np.random.default_rng(303)
    A = rng.uniform(-1, 1, (d, d))
    Bm = rng.uniform(-1, 1, (d, d))
    # one ciphertext per entry; the slots would carry independent matrices: slot s holds entry * (1 + (s % 4))
    ramp = 1.0 + (np.arange(n // 2) % 4)
    enc = lambda val, seed: S.encrypt(val * ramp, sc, L, seed=seed)
    cta = np.empty((d * d, 2, L, n), dtype=np.uint64)
    ctb = np.empty((d * d, 2, L, n), dtype=np.uint64)
    for j in range(d):
        for i in range(d):
            cta[i + j * d] = enc(A[i, j], i + j * d)              # column-major (he_linalg.cpp:376-379)
            ctb[i + j * d] = enc(Bm[i, j], 5000 + i + j * d)
    ctx = hg.Context(n, S.moduli)
    ctx.load_relin_key(S.rk)
    TA = ctx.upload_ct(cta, sc, size_cap=2, L_cap=L)
    TB = ctx.upload_ct(ctb, sc, size_cap=2, L_cap=L)
    out = ctx.ct(d * d, 2, L)
    ctx.matmul_elemwise(out, TA, TB, d, d, d)
    assert out.L == L - 1
    C = A @ Bm
    o = S.o
    worst = 0.0
    for (i, j) in ((0, 0), (63, 63), (17, 42), (5, 60)):
        acc = None
        for k in range(d):
            t = o.multiply(cta[i + k * d], ctb[k + j * d])
            acc = t if acc is None else o.add(acc, t)
        want = o.rescale(o.relinearize(acc, S.rk))
        got = out.download_one(i + j * d)
        assert np.array_equal(got, want), (i, j)
        dec = S.decrypt(got, out.scale).real
        worst = max(worst, float(np.max(np.abs(dec - C[i, j] * ramp * ramp))))
    print(f"cfg3 Matrix::matmul 64x64: max |err| = {worst:.3e}")
    assert worst < 3e-7  # measured 3.0e-8 (values up to 16, one relinearisation per output)


def test_shipped_bfft_n128(tmp_path):
    """The shipped bfft demo (fft.cpp:127-241): N = 16384, {60, 31, 30 x 9, 60}, scale 2^30, n = 128, data[i] = i + 7.1
    replicated over the slots, default (power-of-two) Galois keys; 7 stages.  Through the C++ host mirror, replayed
    on the oracle with the stage plaintexts the host encoded; decrypted result against numpy's FFT (bit-reversed)."""
    n, m = 16384, 128
    S = setup(n, (60, 31) + (30,) * 9 + (60,))
    sc, L = 2.0**30, 11
    data = np.arange(m) + 7.1
    ct = S.encrypt(np.tile(data, n // 2 // m), sc, L, seed=3)
    steps = [sg * (1 << k) for k in range(7) for sg in (1, -1)]
    outs, pts, gk, _ = run(tmp_path, S, "bfft", [m, 0], cts=[(ct, sc)], gk_steps=steps)
    o = S.o
    y, pi = ct, 0
    for i in range(1, 8):
        st, with_d2 = m >> i, i != 1
        y0 = o.rescale(o.multiply_plain(y, pts[pi][0]))
        y1 = o.rescale(o.multiply_plain(o.rotate(y, st, gk)[0], pts[pi + 1][0]))
        nxt = o.add(y0, y1)
        if with_d2:
            nxt = o.add(nxt, o.rescale(o.multiply_plain(o.rotate(y, -st, gk)[0], pts[pi + 2][0])))
        pi += 3 if with_d2 else 2
        y = nxt
    got, got_scale = outs[0]
    assert got.shape == (2, L - 7, n)
    assert np.array_equal(got, y)
    dec = S.decrypt(got, got_scale)[:m]
    ref = np.fft.fft(data)
    brev = np.array([int(format(i, "07b")[::-1], 2) for i in range(m)])
    err = float(np.max(np.abs(dec - ref[brev])))
    print(f"shipped bfft: max |err| = {err:.3e} on outputs up to {np.max(np.abs(ref)):.1f}")
    assert err < 5e-6 * float(np.max(np.abs(ref)))  # measured 5.0e-3 on outputs up to 9.0e3 (30-bit scale, 7 levels): 5.5e-7 relative


def test_shipped_fft_128_ciphertexts(tmp_path):
    """The shipped fft demo (fft.cpp:13-125): N = 16384, {60, 30 x 10, 60}, scale 2^30, 128 ciphertexts with
    vec[i][s] = 2.2 i - 10.8 s + 513.1 in every slot s; recursive radix-2 over the vector of ciphertexts, 448
    butterflies, 7 levels.  Through the C++ host mirror; replayed on the oracle with the host's plaintext limbs."""
    n, cnt = 16384, 128
    S = setup(n, (60,) + (30,) * 10 + (60,))
    sc, L = 2.0**30, 11
    slots = n // 2
    sidx = np.arange(slots)
    vals = np.array([2.2 * i - 10.8 * sidx + 513.1 for i in range(cnt)])  # [cnt][slots]
    cts = [S.encrypt(vals[i], sc, L, seed=i) for i in range(cnt)]
    outs, pts, _, _ = run(tmp_path, S, "fft", [0], cts=[(c, sc) for c in cts])
    assert len(outs) == cnt
    o = S.o
    log = iter(pts)

    def rec(v):
        if len(v) == 1:
            return v
        e, od = rec(v[0::2]), rec(v[1::2])
        one = next(log)[0]
        half = len(e)
        ws = [next(log)[0] for _ in range(half)]
        top, bot = [], []
        for k in range(half):
            t = o.rescale(o.multiply_plain(od[k], ws[k]))
            ee = o.rescale(o.multiply_plain(e[k], one))
            top.append(o.add(ee, t))
            bot.append(o.sub(ee, t))
        return top + bot

    want = rec(cts)
    for idx, ((got, _), w) in enumerate(zip(outs, want)):
        assert got.shape == (2, L - 7, n)
        assert np.array_equal(got, w), idx
    ref = np.fft.fft(vals, axis=0)  # over the ciphertext index, independently per slot
    worst = 0.0
    for k in (0, 1, 64, 127):
        dec = S.decrypt(outs[k][0], outs[k][1])
        worst = max(worst, float(np.max(np.abs(dec - ref[k]))))
    print(f"shipped fft: max |err| = {worst:.3e} on outputs up to {np.max(np.abs(ref)):.3e}")
    assert worst < 5e-10 * float(np.max(np.abs(ref)))  # measured 5.1e-4 on outputs up to 1.1e7


def test_matmul_pow(tmp_path):
    """Matrix::matmul_pow (he_linalg.cpp:316-349), square-and-multiply with the least significant bit first.
    Power 4 = two matmul_square calls: bit-exact against the oracle.  Power 3 multiplies A (level L) by A^2
    (level L-1): SEAL rejects that with "encrypted1 and encrypted2 parameter mismatch" -- the reference's
    own matpow demo is BFV -- and so does the mirror, with SEAL's exception type and message."""
    n, d = 8192, 2
    S = setup(n, (60, 40, 40, 60))
    L, sc = 3, 2.0**40
    rng = np.random.default_rng(404)
    A = rng.uniform(-1, 1, (d, d))
    cts = [S.encrypt(np.full(n // 2, A[i, j]), sc, L, seed=i + j * d) for j in range(d) for i in range(d)]  # column-major
    outs, _, _, _ = run(tmp_path, S, "matpow", [d, 4], cts=[(c, sc) for c in cts], rk=True)
    o = S.o

    def square(m):  # C(i,j) = relin_rescale(sum_k m(i,k) m(k,j)), column-major
        res = []
        for j in range(d):
            for i in range(d):
                acc = None
                for k in range(d):
                    t = o.multiply(m[i + k * d], m[k + j * d])
                    acc = t if acc is None else o.add(acc, t)
                res.append(o.rescale(o.relinearize(acc, S.rk)))
        return res

    want = square(square(cts))
    assert len(outs) == d * d
    A4 = np.linalg.matrix_power(A, 4)
    for idx, ((got, got_scale), w) in enumerate(zip(outs, want)):
        assert np.array_equal(got, w), idx
        assert abs(S.decrypt(got, got_scale).real[0] - A4[idx % d, idx // d]) < 1e-4
    _, _, _, stdout = run(tmp_path, S, "matpow", [d, 3], cts=[(c, sc) for c in cts], rk=True)
    assert "matpow_error=encrypted1 and encrypted2 parameter mismatch" in stdout


def test_least_squares_2d_demo(tmp_path):
    """bench_he_least_squares_2d (src/demos/matrix_operations.cpp:833-1040) end to end through the C++ host mirror at
    the demo's own parameters: N = 32768, {60, 40 x 15, 60}, scale 2^40, the demo's five (x, y) points.  sum_elems
    (rotations 1 and 2), BatchedVector square / product, the operator DSL, he::math::signed_inv (6 iterations) and
    he::util::reach_chain_level, all on the device; slot 0 of the results against the closed-form line fit."""
    n = 32768
    S = setup(n, (60,) + (40,) * 15 + (60,))
    sc, L = 2.0**40, 16
    x = np.array([6, 5.8, 6.5, 5.4, 6.8])
    y = np.array([2, 1.4, 2.4, 1.5, 2.4])
    cx, cy = S.encrypt(x, sc, L, seed=1), S.encrypt(y, sc, L, seed=2)
    outs, _, _, _ = run(tmp_path, S, "least_squares", [len(x)], cts=[(cx, sc), (cy, sc)], rk=True, gk_steps=[1, 2])
    k = len(x)
    denom = k * np.sum(x * x) - np.sum(x) ** 2
    a_num = k * np.sum(x * y) - np.sum(x) * np.sum(y)
    b_num = np.sum(y) * np.sum(x * x) - np.sum(x) * np.sum(x * y)
    want = [denom, 1.0 / denom, a_num, b_num, a_num / denom, b_num / denom]
    got = [float(S.decrypt(ct, scale).real[0]) for ct, scale in outs]
    slope, intercept = np.polyfit(x, y, 1)
    assert abs(want[4] - slope) < 1e-9 and abs(want[5] - intercept) < 1e-9
    for name, g, w in zip(("denom", "1/denom", "a_num", "b_num", "a", "b"), got, want):
        assert abs(g - w) < 2e-4 * max(1.0, abs(w)), (name, g, w)
