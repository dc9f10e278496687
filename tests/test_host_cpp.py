"""The C++ host mirror of the reference's interface (he_operators / he_linalg / he_fft / he_util
over he::gpu types, `homomorphic-encryption-algorithms-diploma-thesis_b200/host/`) driven
through its test binary on the GPU and replayed call by call on the CPU oracle: every result
must be bit-identical.  Routines that encode plaintexts on the host (bfft, fft, drop_chain_levels)
dump the plaintext limbs they produced, and the replay uses exactly those limbs (encode is an
FP-order problem); their decrypted results are also checked against numpy."""
import os
import struct
import subprocess

import numpy as np
import pytest

import hegpu_loader
from fixtures import ckks_tol, rand_residues, setup
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

BIN = os.path.join(hegpu_loader.PKG_DIR, "he_host_test")


def write_case(path, S, cts, pts=(), rk=False, gk_steps=()):
    """cts: list of (array [size][L][N], scale); pts: list of (array [L][N], scale)"""
    with open(path, "wb") as f:
        f.write(struct.pack("<II", S.n, S.K))
        f.write(np.array(S.moduli, dtype=np.uint64).tobytes())
        f.write(struct.pack("<I", 1 if rk else 0))
        if rk:
            f.write(np.ascontiguousarray(S.rk).tobytes())
        gk = S.gk(gk_steps)
        f.write(struct.pack("<I", len(gk)))
        for elt, key in gk.items():
            f.write(struct.pack("<I", elt))
            f.write(np.ascontiguousarray(key).tobytes())
        f.write(struct.pack("<I", len(cts)))
        for a, sc in cts:
            f.write(struct.pack("<IId", a.shape[0], a.shape[1], sc))
            f.write(np.ascontiguousarray(a).tobytes())
        f.write(struct.pack("<I", len(pts)))
        for a, sc in pts:
            f.write(struct.pack("<Id", a.shape[0], sc))
            f.write(np.ascontiguousarray(a).tobytes())
    return gk


def read_out(path, n):
    buf = open(path, "rb").read()
    off = 0
    cts, pts = [], []
    (cnt,) = struct.unpack_from("<I", buf, off)
    off += 4
    for _ in range(cnt):
        size, L, sc = struct.unpack_from("<IId", buf, off)
        off += 16
        a = np.frombuffer(buf, dtype=np.uint64, count=size * L * n, offset=off).reshape(size, L, n).copy()
        off += a.nbytes
        cts.append((a, sc))
    (cnt,) = struct.unpack_from("<I", buf, off)
    off += 4
    for _ in range(cnt):
        L, sc = struct.unpack_from("<Id", buf, off)
        off += 12
        a = np.frombuffer(buf, dtype=np.uint64, count=L * n, offset=off).reshape(L, n).copy()
        off += a.nbytes
        pts.append((a, sc))
    return cts, pts


def run(tmp_path, S, cmd, args, **case):
    assert os.path.exists(BIN), "host mirror not built: run __graft_entry__.build()"
    cpath, opath = str(tmp_path / "case.bin"), str(tmp_path / "out.bin")
    gk = write_case(cpath, S, **case)
    r = subprocess.run([BIN, cpath, cmd, opath] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    cts, pts = read_out(opath, S.n)
    return cts, pts, gk, r.stdout


def test_operator_dsl(tmp_path):
    S = setup(8192, (60, 40, 40, 60))
    rng = np.random.default_rng(1)
    L, sc = 3, 2.0**40
    a, b = rand_residues(rng, S.moduli[:L], (2,), S.n), rand_residues(rng, S.moduli[:L], (2,), S.n)
    p = rand_residues(rng, S.moduli[:L], (), S.n)
    outs, _, gk, _ = run(tmp_path, S, "operators", [], cts=[(a, sc), (b, sc)], pts=[(p, sc)], rk=True, gk_steps=[1, 2, 4, -1, -2, -4])
    o = S.o
    prod = o.multiply(a, b)
    rel = o.relinearize(prod, S.rk)
    want = [o.negate(a)] * 2 + [o.add(a, b)] * 2 + [o.add_plain(a, p)] * 2 + [o.sub(a, b)] * 2 + [o.sub_plain(a, p)] * 2
    want += [prod] * 2 + [o.multiply_plain(a, p)] * 2 + [rel] * 2 + [o.rescale(rel)] * 2 + [o.mod_switch(a)] * 2
    want += [o.rotate(a, 3, gk)[0], o.rotate(a, 1, gk)[0], o.rotate(a, -2, gk)[0], o.rotate(a, -5, gk)[0]]
    assert len(outs) == len(want) == 24
    for i, ((got, _), w) in enumerate(zip(outs, want)):
        assert np.array_equal(got, w), f"operator #{i}"
    assert outs[10][1] == sc * sc and outs[16][1] == sc * sc / S.moduli[2]


def test_errors_keep_seal_types_and_messages(tmp_path):
    S = setup(8192, (60, 40, 40, 60))
    rng = np.random.default_rng(2)
    a = rand_residues(rng, S.moduli[:3], (2,), S.n)
    _, _, _, stdout = run(tmp_path, S, "errors", [], cts=[(a, 2.0**40), (a, 2.0**40)])
    assert "errors_ok=5" in stdout


@pytest.mark.parametrize("case_b", [0, 1])
def test_batched_matrix_matmul(tmp_path, case_b):
    n, dim = 8192, 4
    S = setup(n, (60, 40, 40, 60))
    rng = np.random.default_rng(3)
    L, sc = 3, 2.0**40
    this_c = rand_residues(rng, S.moduli[:L], (dim, 2), n)
    other_c = rand_residues(rng, S.moduli[:L], (dim, 2), n)
    cts = [(x, sc) for x in this_c] + [(x, sc) for x in other_c]
    outs, _, gk, _ = run(tmp_path, S, "bmatmul", [case_b, dim, dim, dim], cts=cts, rk=True, gk_steps=[1, 2, 4, -1, -2, -4])
    o = S.o
    for i in range(dim):
        acc = None
        for j in range(dim):
            src, st = (other_c[i], j) if not case_b else (other_c[j], i)
            t = o.multiply(o.rotate(src, st, gk)[0], this_c[j])
            acc = t if acc is None else o.add(acc, t)
        assert np.array_equal(outs[i][0], o.rescale(o.relinearize(acc, S.rk)))


@pytest.mark.parametrize("rows,inner,cols,at,bt", [(2, 2, 2, 0, 0), (2, 3, 2, 1, 1), (3, 2, 1, 0, 1)])
def test_matrix_matmul_family(tmp_path, rows, inner, cols, at, bt):
    n = 4096
    S = setup(n, (36, 36, 37))
    rng = np.random.default_rng(4)
    L, sc = 2, 2.0**15
    A = rand_residues(rng, S.moduli[:L], (rows * inner, 2), n)  # physical (stored) order
    B = rand_residues(rng, S.moduli[:L], (inner * cols, 2), n)
    outs, _, _, _ = run(tmp_path, S, "matmul", [rows, inner, cols, at, bt], cts=[(x, sc) for x in A] + [(x, sc) for x in B], rk=True)
    o = S.o
    ia = (lambda i, k: i * inner + k) if at else (lambda i, k: i + k * rows)
    ib = (lambda k, j: k * cols + j) if bt else (lambda k, j: k + j * inner)

    def dot(terms):
        acc = None
        for x, y in terms:
            t = o.multiply(x, y)
            acc = t if acc is None else o.add(acc, t)
        return o.rescale(o.relinearize(acc, S.rk))

    want = [dot([(A[ia(i, k)], B[ib(k, j)]) for k in range(inner)]) for j in range(cols) for i in range(rows)]
    if rows == inner == cols and not at:
        d = rows
        want += [dot([(A[ia(i, k)], A[ia(k, j)]) for k in range(d)]) for j in range(d) for i in range(d)]  # A A
        want += [dot([(A[ia(k, i)], A[ia(k, j)]) for k in range(d)]) for j in range(d) for i in range(d)]  # A^T A
        want += [dot([(A[i + j * d], B[i + j * d])]) for j in range(d) for i in range(d)]                    # element-wise
    assert len(outs) == len(want)
    for idx, ((got, _), w) in enumerate(zip(outs, want)):
        assert np.array_equal(got, w), idx


def test_sum_elems_and_square(tmp_path):
    n = 8192
    S = setup(n, (60, 40, 40, 60))
    vals = np.array([-11, 8, 8, 7, -10, 80, 4, 2, 3, 1], dtype=float)  # matrix_operations.cpp:751-757 (expects 92)
    sc, L, dim = 2.0**40, 3, len(vals)
    ct = S.encrypt(vals, sc, L, seed=7)
    outs, _, gk, _ = run(tmp_path, S, "sum_elems", [dim], cts=[(ct, sc)], rk=True, gk_steps=[1, 2, 4, 8])
    o = S.o
    rot = lambda c, s: o.rotate(c, s, gk)[0]
    # replay of he_linalg.cpp:667-713 on the oracle
    bvec, rest, have, window = ct, ct, False, 1
    if dim & 1:
        have, rest = True, rot(rest, window)
    bits = dim >> 1
    while bits:
        window <<= 1
        if bits & 1:
            steps = window >> 1
            acc = o.add(rest, rot(rest, steps))
            steps >>= 1
            while steps:
                acc = o.add(acc, rot(acc, steps))
                steps >>= 1
            bvec, have = (o.add(bvec, acc), True) if have else (acc, True)
            if bits != 1:
                rest = rot(rest, window)
        bits >>= 1
    assert np.array_equal(outs[0][0], bvec)
    assert abs(S.decrypt(outs[0][0], sc).real[0] - 92.0) < ckks_tol(dim, n, sc) * 100
    assert np.array_equal(outs[1][0], o.rescale(o.relinearize(o.square(ct), S.rk)))


@pytest.mark.parametrize("inverse", [0, 1])
def test_bfft(tmp_path, inverse):
    n, m = 8192, 8
    S = setup(n, (60, 40, 40, 40, 40, 60))
    sc, L = 2.0**40, 5
    data = np.arange(m) + 7.1  # fft.cpp:163-177
    ct = S.encrypt(np.tile(data, n // 2 // m), sc, L, seed=3)
    outs, pts, gk, _ = run(tmp_path, S, "bfft", [m, inverse], cts=[(ct, sc)], gk_steps=[4, -4, 2, -2, 1, -1])
    o = S.o
    y, pi = ct, 0
    for i in range(1, 4):
        steps, with_d2 = m >> i, i != 1
        d0, d1 = pts[pi][0], pts[pi + 1][0]
        y0 = o.rescale(o.multiply_plain(y, d0))
        y1 = o.rescale(o.multiply_plain(o.rotate(y, steps, gk)[0], d1))
        nxt = o.add(y0, y1)
        if with_d2:
            nxt = o.add(nxt, o.rescale(o.multiply_plain(o.rotate(y, -steps, gk)[0], pts[pi + 2][0])))
        pi += 3 if with_d2 else 2
        y = nxt
    if inverse:
        y = o.rescale(o.multiply_plain(y, pts[pi][0]))
    got, got_scale = outs[0]
    assert np.array_equal(got, y)
    dec = S.decrypt(got, got_scale)[:m]
    ref = np.fft.ifft(data) if inverse else np.fft.fft(data)
    brev = [int(format(i, "03b")[::-1], 2) for i in range(m)]
    assert np.max(np.abs(dec - ref[brev])) < 1e-3  # output is in bit-reversed order (fft.cpp:224-238)


@pytest.mark.parametrize("inverse", [0, 1])
def test_fft_over_ciphertexts(tmp_path, inverse):
    n, cnt = 8192, 4
    S = setup(n, (60, 40, 40, 40, 60))
    sc, L = 2.0**40, 4
    vals = np.array([2.2 * i + 513.1 for i in range(cnt)])  # fft.cpp:19-23 pattern, slot 0
    cts = [S.encrypt(np.full(n // 2, v), sc, L, seed=i) for i, v in enumerate(vals)]
    outs, pts, _, _ = run(tmp_path, S, "fft", [inverse], cts=[(c, sc) for c in cts])
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
    if inverse:
        ninv = next(log)[0]
        want = [o.rescale(o.multiply_plain(w, ninv)) for w in want]
    for (got, _), w in zip(outs, want):
        assert np.array_equal(got, w)
    dec = np.array([S.decrypt(g, s_).real[0] for g, s_ in outs])
    ref = np.fft.ifft(vals) if inverse else np.fft.fft(vals)
    assert np.max(np.abs(dec - ref.real)) < 1e-2


def test_drop_chain_levels(tmp_path):
    n = 8192
    S = setup(n, (60, 40, 40, 60))
    rng = np.random.default_rng(8)
    ct = rand_residues(rng, S.moduli[:3], (2,), n)
    outs, pts, _, _ = run(tmp_path, S, "drop_levels", [1], cts=[(ct, 2.0**40)])
    want = S.o.rescale(S.o.multiply_plain(ct, pts[-1][0]))
    assert np.array_equal(outs[0][0], want) and np.array_equal(outs[1][0], want)
    # SEAL's real-scalar encode: round(1 * scale) in every slot of every limb
    for i, q in enumerate(S.moduli[:3]):
        assert np.all(pts[-1][0][i] == np.uint64(2**40 % q))


class _Replay:
    """The reference's he_math.cpp statement by statement on the CPU oracle (an independent restatement:
    written from src/core/he_math.cpp:22-269, not from host/he_math.cpp).  A value is (ct, scale)."""

    def __init__(self, S):
        self.S, self.o, self.q = S, S.o, S.moduli

    def const(self, value, v):
        ct, sc = v
        return self.S.enc.encode_scalar(value, sc, ct.shape[1])

    def mul_const_rescale(self, v, value):
        ct, sc = v
        out = self.o.rescale(self.o.multiply_plain(ct, self.const(value, v)))
        return out, sc * sc / self.q[ct.shape[1] - 1]

    def add_const(self, v, value):
        return self.o.add_plain(v[0], self.const(value, v)), v[1]

    def sub_const(self, v, value):
        return self.o.sub_plain(v[0], self.const(value, v)), v[1]

    def mul_relin_rescale(self, a, b):
        ct = self.o.rescale(self.o.relinearize(self.o.multiply(a[0], b[0]), self.S.rk))
        return ct, a[1] * b[1] / self.q[a[0].shape[1] - 1]

    def square_relin_rescale(self, a):
        ct = self.o.rescale(self.o.relinearize(self.o.square(a[0]), self.S.rk))
        return ct, a[1] * a[1] / self.q[a[0].shape[1] - 1]

    def signed_inv(self, x, a, iters):  # he_math.cpp:22-90
        y = self.add_const(self.mul_const_rescale(x, -a * a), 2 * a)
        if iters == 1:
            return y
        t = self.mul_const_rescale(x, a)
        one = self.const(1, t)
        t = (self.o.sub_plain(t[0], one), t[1])
        y = (self.o.rescale(self.o.multiply_plain(y[0], one)), y[1] * t[1] / self.q[y[0].shape[1] - 1])
        for _ in range(1, iters):
            t = self.square_relin_rescale(t)
            y = self.mul_relin_rescale(y, self.add_const(t, 1))
        return y

    def inv_sqrt_twice(self, x, a, iters):  # he_math.cpp:95-160 (the compiled "#if 1" branch)
        y = self.add_const(self.mul_const_rescale(x, -a * a * a), 1.5 * a)
        for i in range(1, iters):
            yp = y
            y = self.mul_const_rescale(self.mul_const_rescale(y, 1.5), 1)
            for _ in range(2 if i > 1 else 1):
                x = self.mul_const_rescale(x, 1)
            xy = self.mul_relin_rescale(x, yp)
            yp = self.mul_relin_rescale(self.square_relin_rescale(yp), xy)
            y = (self.o.sub(y[0], yp[0]), y[1])
        return y

    def sqrt(self, x, a, iters):  # he_math.cpp:210-232
        y = self.inv_sqrt_twice(x, 1 / a / np.sqrt(2), iters)
        sx = self.mul_const_rescale(x, np.sqrt(2))
        while sx[0].shape[1] > y[0].shape[1]:
            sx = self.mul_const_rescale(sx, 1)
        return self.mul_relin_rescale(y, sx)

    def abs(self, x, a, iters):  # he_math.cpp:237-269
        x2 = self.square_relin_rescale(x)
        y = self.inv_sqrt_twice(x2, 1 / a / np.sqrt(2), iters)
        x2 = self.mul_const_rescale(x2, np.sqrt(2))
        while x2[0].shape[1] > y[0].shape[1]:
            x2 = self.mul_const_rescale(x2, 1)
        return self.mul_relin_rescale(y, x2)


@pytest.mark.parametrize("which,name,iters", [(0, "signed_inv", 4), (0, "signed_inv", 1), (1, "inv_sqrt_twice", 3), (2, "sqrt", 3), (3, "abs", 3)])
def test_he_math(tmp_path, which, name, iters):
    """he::math through the C++ host mirror: bit-identical to the reference's statement
    sequence replayed on the oracle, and close to the real function after decryption."""
    n = 8192
    S = setup(n, (60,) + (40,) * 7 + (60,))
    sc, L, a = 2.0**40, 8, 1.0
    rng = np.random.default_rng(70 + which)
    slots = n // 2
    if name == "signed_inv":
        x = rng.uniform(0.6, 1.4, slots)
        ref = 1 / x
    elif name == "inv_sqrt_twice":
        x = rng.uniform(0.4, 0.6, slots)
        ref = 1 / np.sqrt(2 * x)
    elif name == "sqrt":
        x = rng.uniform(0.8, 1.2, slots)
        ref = np.sqrt(x)
    else:
        x = rng.uniform(0.8, 1.2, slots) * rng.choice([-1.0, 1.0], slots)
        ref = np.abs(x)
    ct = S.encrypt(x, sc, L, seed=which)
    outs, _, _, _ = run(tmp_path, S, "math", [which, a, iters], cts=[(ct, sc)], rk=True)
    want, want_scale = getattr(_Replay(S), name)((ct, sc), a, iters)
    got, got_scale = outs[0]
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    assert abs(got_scale / want_scale - 1) < 1e-12
    dec = S.decrypt(got, got_scale).real
    # truncation error of the iteration itself dominates: 0.4^(2^iters) for the product form, quadratic Newton otherwise
    tol = {("signed_inv", 4): 1e-5, ("signed_inv", 1): 0.3, ("inv_sqrt_twice", 3): 2e-3, ("sqrt", 3): 2e-3, ("abs", 3): 5e-3}[(name, iters)]
    assert np.max(np.abs(dec - ref)) < tol
