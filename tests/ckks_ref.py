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


# ---------------------------------------------------------------- spec 9.6 / 9.7 in big ints
def switch_key_ref(ct, target, key, moduli, psis, L):
    """SURVEY 9.6. ct [2][L][N], target [L][N], key [Lmax][2][K][N] as nested int lists."""
    K = len(moduli)
    P = moduli[K - 1]
    half = P // 2
    n = len(target[0])
    coef = [intt_naive(target[j], moduli[j], psis[j]) for j in range(L)]
    acc = [[None] * (L + 1) for _ in range(2)]
    for I in range(L + 1):
        ki = K - 1 if I == L else I
        m = moduli[ki]
        ops = []
        for J in range(L):
            if I == J:
                ops.append(target[J])
            else:
                ops.append(ntt_naive([x % m for x in coef[J]], m, psis[ki]))
        for c in range(2):
            acc[c][I] = [sum(ops[J][x] * key[J][c][ki][x] for J in range(L)) % m for x in range(n)]
    out = [[None] * L for _ in range(2)]
    for c in range(2):
        t = intt_naive(acc[c][L], P, psis[K - 1])
        t = [(x + half) % P for x in t]
        for i in range(L):
            q = moduli[i]
            d = ntt_naive([(x % q - half % q) % q for x in t], q, psis[i])
            pinv = pow(P % q, q - 2, q)
            out[c][i] = [(ct[c][i][x] + (acc[c][i][x] - d[x]) * pinv) % q for x in range(n)]
    return out


def rescale_ref(ct, moduli, psis, L):
    """SURVEY 9.7. ct [size][L][N] -> [size][L-1][N]."""
    ql = moduli[L - 1]
    half = ql // 2
    out = []
    for poly in ct:
        t = intt_naive(poly[L - 1], ql, psis[L - 1])
        t = [(x + half) % ql for x in t]
        res = []
        for i in range(L - 1):
            q = moduli[i]
            d = ntt_naive([(x % q - half % q) % q for x in t], q, psis[i])
            inv = pow(ql % q, q - 2, q)
            res.append([(poly[i][x] - d[x]) * inv % q for x in range(len(t))])
        out.append(res)
    return out


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
