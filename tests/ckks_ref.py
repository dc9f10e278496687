"""Independent pure-Python restatement used to cross-check the C oracle and the CUDA path.

Test infrastructure only.  Two parts:

* big-integer restatements of SURVEY.md 9.2 / 9.6 / 9.7 (naive NTT by direct evaluation,
  switch_key, rescale) written from the spec, *not* from oracle/ckks_oracle.c -- slow,
  for small N only.
* CKKS encode / decode in numpy + Python ints, used to build plaintexts
  and to check decrypted results against float64 linear algebra.
"""
from __future__ import annotations

import numpy as np


# ---------------------------------------------------------------- small helpers
def brev(x: int, bits: int) -> int:
    r = 0
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r


def is_prime(n: int) -> bool:
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, r = n - 1, 0
    while d % 2 == 0:
        d //= 2
        r += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(r - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def get_primes(factor: int, bits: int, count: int) -> list[int]:
    """SEAL util::get_primes."""
    v = ((1 << bits) - 1) // factor * factor + 1
    lo = 1 << (bits - 1)
    out = []
    while len(out) < count and v > lo:
        if is_prime(v):
            out.append(v)
        v -= factor
    assert len(out) == count
    return out


def coeff_modulus_create(n: int, bits: list[int]) -> list[int]:
    """SEAL CoeffModulus::Create: smallest prime of each size first."""
    table = {b: get_primes(2 * n, b, bits.count(b)) for b in set(bits)}
    return [table[b].pop() for b in bits]


def minimal_primitive_root(q: int, n: int) -> int:
    e = (q - 1) // (2 * n)
    g = 2
    while True:
        r = pow(g, e, q)
        if pow(r, n, q) == q - 1:
            break
        g += 1
    best, cur, sq = r, r, r * r % q
    for _ in range(n):
        best = min(best, cur)
        cur = cur * sq % q
    return best


# ---------------------------------------------------------------- big-int NTT (spec 9.2)
def ntt_naive(a: list[int], q: int, psi: int) -> list[int]:
    """a^[k] = a(psi^(2 brev(k)+1)) by direct evaluation: O(N^2)."""
    n = len(a)
    logn = n.bit_length() - 1
    out = []
    for k in range(n):
        x = pow(psi, 2 * brev(k, logn) + 1, q)
        acc = 0
        for c in reversed(a):
            acc = (acc * x + c) % q
        out.append(acc)
    return out


def intt_naive(ah: list[int], q: int, psi: int) -> list[int]:
    n = len(ah)
    logn = n.bit_length() - 1
    ninv = pow(n, q - 2, q)
    ipsi = pow(psi, q - 2, q)
    out = []
    for j in range(n):
        acc = 0
        for k in range(n):
            acc += ah[k] * pow(ipsi, (2 * brev(k, logn) + 1) * j, q)
        out.append(acc % q * ninv % q)
    return out


def galois_coeff(a: list[int], elt: int, q: int) -> list[int]:
    """Coefficient-domain automorphism X -> X^elt in Z_q[X]/(X^N+1)."""
    n = len(a)
    out = [0] * n
    for i, c in enumerate(a):
        e = i * elt % (2 * n)
        if e < n:
            out[e] = c % q
        else:
            out[e - n] = (-c) % q
    return out


# ---------------------------------------------------------------- big-int NTT in O(N log N)
def _dft(b: list[int], w: int, q: int) -> list[int]:
    """Cyclic DFT over Z_q, natural order in and out: recursive radix-2 decimation in time (textbook form,
    nothing shared with the oracle's in-place Harvey butterflies)."""
    n = len(b)
    if n == 1:
        return b
    w2 = w * w % q
    ev, od = _dft(b[0::2], w2, q), _dft(b[1::2], w2, q)
    out = [0] * n
    t, h = 1, n // 2
    for j in range(h):
        x = t * od[j] % q
        out[j] = (ev[j] + x) % q
        out[j + h] = (ev[j] - x) % q
        t = t * w % q
    return out


def ntt_fast(a: list[int], q: int, psi: int) -> list[int]:
    """Same function as ntt_naive (a^[k] = a(psi^(2 brev(k)+1))): twist by psi^i, cyclic DFT with omega = psi^2,
    bit-reversed read-out.  Checked against ntt_naive in tests/test_oracle_evaluator.py."""
    n = len(a)
    logn = n.bit_length() - 1
    tw, p = [], 1
    for c in a:
        tw.append(c * p % q)
        p = p * psi % q
    big = _dft(tw, psi * psi % q, q)
    return [big[brev(k, logn)] for k in range(n)]


def intt_fast(ah: list[int], q: int, psi: int) -> list[int]:
    n = len(ah)
    logn = n.bit_length() - 1
    nat = [0] * n
    for k in range(n):
        nat[brev(k, logn)] = ah[k]
    ipsi = pow(psi, q - 2, q)
    b = _dft(nat, ipsi * ipsi % q, q)
    ninv = pow(n, q - 2, q)
    out, p = [], ninv
    for c in b:
        out.append(c * p % q)
        p = p * ipsi % q
    return out


# the spec restatements below call the transforms through these two names: O(N^2) direct evaluation by default,
# the O(N log N) form inside `with fast_transforms():` (full-size cross-checks)
_NTT, _INTT = ntt_naive, intt_naive


class fast_transforms:
    def __enter__(self):
        global _NTT, _INTT
        _NTT, _INTT = ntt_fast, intt_fast

    def __exit__(self, *exc):
        global _NTT, _INTT
        _NTT, _INTT = ntt_naive, intt_naive
        return False


def apply_galois_ref(ct, elt, key, moduli, psis, L):
    """SURVEY 9.4 apply_galois_inplace on a size-2 ciphertext [2][L][N]: c0 <- pi(c0); switch_key(pi(c1))."""
    n = len(ct[0][0])
    tab = galois_table_ntt(n, elt)
    c0 = [[ct[0][i][tab[x]] for x in range(n)] for i in range(L)]
    t = [[ct[1][i][tab[x]] for x in range(n)] for i in range(L)]
    zero = [[0] * n for _ in range(L)]
    return switch_key_ref([c0, zero], t, key, moduli, psis, L)


# ---------------------------------------------------------------- spec 9.6 / 9.7 in big ints
def switch_key_ref(ct, target, key, moduli, psis, L):
    """SURVEY 9.6. ct [2][L][N], target [L][N], key [Lmax][2][K][N] as nested int lists."""
    K = len(moduli)
    P = moduli[K - 1]
    half = P // 2
    n = len(target[0])
    coef = [_INTT(target[j], moduli[j], psis[j]) for j in range(L)]
    acc = [[None] * (L + 1) for _ in range(2)]
    for I in range(L + 1):
        ki = K - 1 if I == L else I
        m = moduli[ki]
        ops = []
        for J in range(L):
            if I == J:
                ops.append(target[J])
            else:
                ops.append(_NTT([x % m for x in coef[J]], m, psis[ki]))
        for c in range(2):
            acc[c][I] = [sum(ops[J][x] * key[J][c][ki][x] for J in range(L)) % m for x in range(n)]
    out = [[None] * L for _ in range(2)]
    for c in range(2):
        t = _INTT(acc[c][L], P, psis[K - 1])
        t = [(x + half) % P for x in t]
        for i in range(L):
            q = moduli[i]
            d = _NTT([(x % q - half % q) % q for x in t], q, psis[i])
            pinv = pow(P % q, q - 2, q)
            out[c][i] = [(ct[c][i][x] + (acc[c][i][x] - d[x]) * pinv) % q for x in range(n)]
    return out


def rescale_ref(ct, moduli, psis, L):
    """SURVEY 9.7. ct [size][L][N] -> [size][L-1][N]."""
    ql = moduli[L - 1]
    half = ql // 2
    out = []
    for poly in ct:
        t = _INTT(poly[L - 1], ql, psis[L - 1])
        t = [(x + half) % ql for x in t]
        res = []
        for i in range(L - 1):
            q = moduli[i]
            d = _NTT([(x % q - half % q) % q for x in t], q, psis[i])
            inv = pow(ql % q, q - 2, q)
            res.append([(poly[i][x] - d[x]) * inv % q for x in range(len(t))])
        out.append(res)
    return out


def galois_table_ntt(n: int, elt: int) -> list[int]:
    """result[i] = operand[ brev(((elt*(2*brev(i)+1)) >> 1) & (N-1)) ]."""
    logn = n.bit_length() - 1
    return [brev(((elt * (2 * brev(i, logn) + 1)) >> 1) & (n - 1), logn) for i in range(n)]


def galois_elt_from_step(n: int, step: int) -> int:
    m = 2 * n
    if step == 0:
        return m - 1
    pos = abs(step)
    s = (n >> 1) - pos if step < 0 else pos
    return pow(3, s, m)


def _decompose_ref(target, moduli, psis, L):
    """ext[J][I] = NTT_{m_I}(INTT_{q_J}(target_J) mod m_I); ext[J][J] = target_J."""
    K = len(moduli)
    ext = [[None] * (L + 1) for _ in range(L)]
    for J in range(L):
        coef = _INTT(target[J], moduli[J], psis[J])
        for I in range(L + 1):
            ki = K - 1 if I == L else I
            ext[J][I] = list(target[J]) if I == J else _NTT([x % moduli[ki] for x in coef], moduli[ki], psis[ki])
    return ext


def _inner_ref(ext, key, moduli, L, tab=None):
    K = len(moduli)
    n = len(ext[0][0])
    acc = [[None] * (L + 1) for _ in range(2)]
    for I in range(L + 1):
        ki = K - 1 if I == L else I
        for c in range(2):
            acc[c][I] = [sum(ext[J][I][tab[x] if tab else x] * key[J][c][ki][x] for J in range(L)) % moduli[ki] for x in range(n)]
    return acc


def _mod_down_ref(acc, moduli, psis, L):
    """[2][L+1][N] in the basis q_0..q_{L-1},P -> round(acc / P) as [2][L][N]."""
    K = len(moduli)
    P = moduli[K - 1]
    half = P // 2
    out = [[None] * L for _ in range(2)]
    for c in range(2):
        t = [(x + half) % P for x in _INTT(acc[c][L], P, psis[K - 1])]
        for i in range(L):
            q = moduli[i]
            d = _NTT([(x % q - half % q) % q for x in t], q, psis[i])
            pinv = pow(P % q, q - 2, q)
            out[c][i] = [(acc[c][i][x] - d[x]) * pinv % q for x in range(len(t))]
    return out


def _mod_down_one_ref(acc1, moduli, psis, L):
    """One component [L+1][N] in the basis q_0..q_{L-1},P -> round(acc / P) as [L][N]."""
    K = len(moduli)
    P = moduli[K - 1]
    half = P // 2
    t = [(x + half) % P for x in _INTT(acc1[L], P, psis[K - 1])]
    out = []
    for i in range(L):
        q = moduli[i]
        d = _NTT([(x % q - half % q) % q for x in t], q, psis[i])
        pinv = pow(P % q, q - 2, q)
        out.append([(acc1[i][x] - d[x]) * pinv % q for x in range(len(t))])
    return out


def matvec_dh_ref(ct, n1, n2, ptsx, baby_keys, giant_keys, moduli, psis, L, rescale=True, g_first=0):
    """Double-hoisted BSGS matvec (HEGPU_MATVEC_DH) in big integers, written from the description in
    include/hegpu.h -- not from the C oracle.  ct [2][L][N]; ptsx [n1*n2][L+1][N]; keys as in switch_key_ref."""
    K = len(moduli)
    P = moduli[K - 1]
    n = len(ct[0][0])
    mods = [moduli[i] for i in range(L)] + [P]
    ext = _decompose_ref(ct[1], moduli, psis, L)
    baby = []
    for k in range(n1):
        if k == 0:
            b = [[[ct[c][i][x] * P % mods[i] for x in range(n)] for i in range(L)] + [[0] * n] for c in range(2)]
        else:
            tab = galois_table_ntt(n, galois_elt_from_step(n, k))
            b = _inner_ref(ext, baby_keys[k], moduli, L, tab)
            for i in range(L):
                b[0][i] = [(b[0][i][x] + P * ct[0][i][tab[x]]) % mods[i] for x in range(n)]
        baby.append(b)
    F = [[[0] * n for _ in range(L + 1)] for _ in range(2)]
    for g in range(n2):
        u = [[[sum(baby[k][c][I][x] * ptsx[g * n1 + k][I][x] for k in range(n1)) % mods[I] for x in range(n)]
              for I in range(L + 1)] for c in range(2)]
        if g_first + g:
            # only the component that is key-switched leaves the extended basis; c0 is permuted limb-wise and stays
            tab = galois_table_ntt(n, galois_elt_from_step(n, (g_first + g) * n1))
            v1 = _mod_down_one_ref(u[1], moduli, psis, L)
            tgt = [[v1[i][tab[x]] for x in range(n)] for i in range(L)]
            ks = _inner_ref(_decompose_ref(tgt, moduli, psis, L), giant_keys[g], moduli, L)
            add = [[[(u[0][I][tab[x]] + ks[0][I][x]) % mods[I] for x in range(n)] for I in range(L + 1)], ks[1]]
        else:
            add = u
        for c in range(2):
            for I in range(L + 1):
                F[c][I] = [(F[c][I][x] + add[c][I][x]) % mods[I] for x in range(n)]
    res = _mod_down_ref(F, moduli, psis, L)
    return rescale_ref(res, moduli, psis, L) if rescale else res


# ---------------------------------------------------------------- CKKS encode / decode (9.8)
class Encoder:
    """Vector CKKS encoder over ring degree n (slots = n/2). NTTs are delegated to `ntt_fwd(i, a)`
    / `ntt_inv(i, a)` callables (the oracle's in CPU tests)."""

    def __init__(self, n: int, moduli, ntt_fwd, ntt_inv):
        self.n = n
        self.slots = n // 2
        self.moduli = [int(q) for q in moduli]
        self._fwd, self._inv = ntt_fwd, ntt_inv
        m = 2 * n
        pos = 1
        idx1 = np.empty(self.slots, dtype=np.int64)
        idx2 = np.empty(self.slots, dtype=np.int64)
        for i in range(self.slots):
            idx1[i] = (pos - 1) >> 1
            idx2[i] = (m - pos - 1) >> 1
            pos = pos * 3 % m
        self.idx1, self.idx2 = idx1, idx2
        k = np.arange(n)
        self.zeta_pos = np.exp(1j * np.pi * k / n)
        self.zeta_neg = np.conj(self.zeta_pos)

    def coeffs(self, values, scale: float) -> list[int]:
        """Integer coefficient vector round(scale * embedding^{-1}(values))."""
        z = np.zeros(self.slots, dtype=np.complex128)
        values = np.asarray(values, dtype=np.complex128).ravel()
        assert values.size <= self.slots
        z[: values.size] = values
        v = np.zeros(self.n, dtype=np.complex128)
        v[self.idx1] = z
        v[self.idx2] = np.conj(z)
        m = (np.fft.fft(v) * self.zeta_neg / self.n).real * scale
        r = np.rint(m)
        if np.max(np.abs(r)) < 2**62:
            return [int(x) for x in r.astype(np.int64)]
        return [int(x) for x in r]

    def encode(self, values, scale: float, L: int) -> np.ndarray:
        c = self.coeffs(values, scale)
        return self.encode_coeffs(c, L)

    def encode_coeffs(self, c, L: int) -> np.ndarray:
        out = np.empty((L, self.n), dtype=np.uint64)
        small = max(abs(x) for x in c) < 2**62
        ca = np.array(c, dtype=np.int64) if small else None
        for i in range(L):
            q = self.moduli[i]
            if small:
                r = np.mod(ca, np.int64(q)).astype(np.uint64)
            else:
                r = np.array([x % q for x in c], dtype=np.uint64)
            out[i] = self._fwd(i, r)
        return out

    def encode_ext(self, values, scale: float, L: int) -> np.ndarray:
        """[L+1][N]: the L data limbs plus a limb mod the special prime (moduli[-1]) -- the plaintext
        layout of the double-hoisted matvec (HEGPU_MATVEC_DH), i.e. an encode at the key level
        restricted to the limbs q_0..q_{L-1}, P."""
        c = self.coeffs(values, scale)
        out = np.empty((L + 1, self.n), dtype=np.uint64)
        out[:L] = self.encode_coeffs(c, L)
        K = len(self.moduli)
        P = self.moduli[K - 1]
        out[L] = self._fwd(K - 1, np.array([x % P for x in c], dtype=np.uint64))
        return out

    def encode_scalar(self, value: float, scale: float, L: int) -> np.ndarray:
        """SEAL encode(double): constant round(value*scale) in every NTT slot."""
        v = int(round(value * scale))
        out = np.empty((L, self.n), dtype=np.uint64)
        for i in range(L):
            out[i, :] = v % self.moduli[i]
        return out

    def decode(self, plain_ntt: np.ndarray, scale: float) -> np.ndarray:
        L = plain_ntt.shape[0]
        res = [self._inv(i, plain_ntt[i]) for i in range(L)]
        mods = self.moduli[:L]
        Q = 1
        for q in mods:
            Q *= q
        # CRT compose with Python ints
        acc = [0] * self.n
        for i, q in enumerate(mods):
            Qi = Q // q
            f = Qi * pow(Qi % q, q - 2, q)
            ri = res[i].tolist()
            for j in range(self.n):
                acc[j] += ri[j] * f
        half = Q // 2
        m = np.empty(self.n, dtype=np.float64)
        for j in range(self.n):
            x = acc[j] % Q
            if x > half:
                x -= Q
            m[j] = float(x)
        v = np.fft.ifft(m * self.zeta_pos) * self.n / scale
        return v[self.idx1]
