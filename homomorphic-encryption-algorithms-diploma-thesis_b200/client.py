"""Host-side CKKS client tooling: parameters, encoder, key generation, encryption, decryption.

In the reference these steps are done by Microsoft SEAL on the client (KeyGenerator,
Encryptor, Decryptor, CKKSEncoder -- e.g. src/demos/matrix_operations.cpp:1057-1121) and are
NOT on the evaluator hot path; with SEAL present its objects are handed to the C ABI
directly.  SEAL is absent from this image, so the bench, the demos and
the multi-GPU driver need a stand-in.  Everything with modular arithmetic in it (NTTs,
polynomial products) runs on the GPU through the C ABI; the host only samples randomness
and runs the double-precision embedding FFT (SEAL also encodes on the host).

Conventions follow SEAL 4.1.
"""
from __future__ import annotations

import numpy as np


# ---------------------------------------------------------------- parameters
def _is_prime(n: int) -> bool:
    if n < 2:
        return False
    small = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)
    for p in small:
        if n % p == 0:
            return n == p
    d, r = n - 1, 0
    while d % 2 == 0:
        d //= 2
        r += 1
    for a in small:
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


def coeff_modulus_create(n: int, bit_sizes) -> list[int]:
    """seal::CoeffModulus::Create(n, bit_sizes)."""
    bit_sizes = list(bit_sizes)
    table = {}
    for b in set(bit_sizes):
        need = bit_sizes.count(b)
        v = ((1 << b) - 1) // (2 * n) * (2 * n) + 1
        found = []
        while len(found) < need and v > (1 << (b - 1)):
            if _is_prime(v):
                found.append(v)
            v -= 2 * n
        if len(found) < need:
            raise ValueError("failed to find enough qualifying primes")
        table[b] = found
    return [table[b].pop() for b in bit_sizes]


def _brev_table(logn: int) -> np.ndarray:
    idx = np.arange(1 << logn, dtype=np.uint64)
    r = np.zeros_like(idx)
    for i in range(logn):
        r |= ((idx >> np.uint64(i)) & np.uint64(1)) << np.uint64(logn - 1 - i)
    return r


class Client:
    """Keys + encoder for one parameter set, bound to a GPU Context for the arithmetic."""

    def __init__(self, ctx, seed: int = 0):
        self.ctx = ctx
        self.n = ctx.n
        self.logn = self.n.bit_length() - 1
        self.moduli = list(ctx.moduli)
        self.K = len(self.moduli)
        self.Lmax = self.K - 1
        self.rng = np.random.default_rng(seed)
        self.slots = self.n // 2
        # CKKSEncoder index map
        m = 2 * self.n
        pos = 1
        i1 = np.empty(self.slots, dtype=np.int64)
        i2 = np.empty(self.slots, dtype=np.int64)
        for i in range(self.slots):
            i1[i] = (pos - 1) >> 1
            i2[i] = (m - pos - 1) >> 1
            pos = pos * 3 % m
        self._i1, self._i2 = i1, i2
        k = np.arange(self.n)
        self._zeta = np.exp(1j * np.pi * k / self.n)
        self._brev = _brev_table(self.logn)
        self.secret = None  # [K][N], NTT form
        self.keygen_secret()

    # ------------------------------------------------------------ sampling
    def _to_residues(self, signed: np.ndarray, limbs: int) -> np.ndarray:
        out = np.empty((limbs, self.n), dtype=np.uint64)
        s = signed.astype(np.int64)
        for i in range(limbs):
            out[i] = np.mod(s, np.int64(self.moduli[i])).astype(np.uint64)
        return out

    def _uniform(self, limbs: int, prefix=()) -> np.ndarray:
        out = np.empty(tuple(prefix) + (limbs, self.n), dtype=np.uint64)
        for i in range(limbs):
            out[..., i, :] = self.rng.integers(0, self.moduli[i], size=tuple(prefix) + (self.n,), dtype=np.uint64)
        return out

    def _noise_ntt(self, limbs: int) -> np.ndarray:
        """centred binomial error (sigma ~ 3.2, SEAL's default sampler), NTT form"""
        e = self.rng.binomial(21, 0.5, self.n) - self.rng.binomial(21, 0.5, self.n)
        r = self._to_residues(e, limbs)
        self.ctx.ntt_forward_host(r, 0, limbs)
        return r

    def keygen_secret(self):
        s = self.rng.integers(-1, 2, self.n)
        r = self._to_residues(s, self.K)
        self.ctx.ntt_forward_host(r, 0, self.K)
        self.secret = r

    # ------------------------------------------------------------ GPU helpers
    def _neg_as_plus_e(self, a: np.ndarray, e: np.ndarray, limbs: int) -> np.ndarray:
        """-(a*s + e) for a batch: a, e [B][limbs][N] -> [B][limbs][N] (GPU dyadic ops)"""
        ctx = self.ctx
        B = a.shape[0]
        two = np.ascontiguousarray(np.stack([a, a], axis=1))  # poly 0 is the operand, poly 1 is ignored
        A = ctx.upload_ct(two, 1.0, size_cap=2, L_cap=limbs)
        S = ctx.upload_pt(np.ascontiguousarray(self.secret[:limbs]), 1.0, L_cap=limbs)
        E = ctx.upload_pt(np.ascontiguousarray(e), 1.0, L_cap=limbs)
        ctx.multiply_plain(A, A, S, 0)
        A.scale = 1.0
        ctx.add_plain(A, A, E, -1)
        ctx.negate(A, A)
        return np.ascontiguousarray(A.download()[:, 0])

    def _scalar_mul(self, poly: np.ndarray, factor: int, q: int) -> np.ndarray:
        return np.array([(int(x) * factor) % q for x in poly], dtype=np.uint64)

    # ------------------------------------------------------------ keys
    def kswitch_key(self, new_key: np.ndarray) -> np.ndarray:
        """KeyGenerator::generate_one_kswitch_key: [Lmax][2][K][N]"""
        K, Lmax, n = self.K, self.Lmax, self.n
        a = self._uniform(K, (Lmax,))
        e = np.stack([self._noise_ntt(K) for _ in range(Lmax)])
        c0 = self._neg_as_plus_e(a, e, K)
        P = self.moduli[K - 1]
        for j in range(Lmax):
            q = self.moduli[j]
            add = self._scalar_mul(new_key[j], P % q, q)
            c0[j, j] = ((c0[j, j].astype(object) + add.astype(object)) % q).astype(np.uint64)
        key = np.empty((Lmax, 2, K, n), dtype=np.uint64)
        key[:, 0] = c0
        key[:, 1] = a
        return key

    def relin_key(self) -> np.ndarray:
        ctx = self.ctx
        s2ct = np.ascontiguousarray(np.stack([self.secret, self.secret])[None])
        A = ctx.upload_ct(s2ct, 1.0, size_cap=2, L_cap=self.K)
        S = ctx.upload_pt(self.secret, 1.0, L_cap=self.K)
        ctx.multiply_plain(A, A, S, 0)
        s2 = np.ascontiguousarray(A.download()[0, 0])
        return self.kswitch_key(s2)

    def galois_table(self, elt: int) -> np.ndarray:
        """GaloisTool::generate_table_ntt"""
        r = 2 * self._brev + 1
        raw = ((np.uint64(elt) * r) >> np.uint64(1)) & np.uint64(self.n - 1)
        return self._brev[raw.astype(np.int64)].astype(np.int64)

    def galois_key(self, elt: int) -> np.ndarray:
        tab = self.galois_table(elt)
        return self.kswitch_key(np.ascontiguousarray(self.secret[:, tab]))

    def galois_keys_for_steps(self, steps) -> dict:
        return {self.ctx.galois_elt_from_step(s): self.galois_key(self.ctx.galois_elt_from_step(s)) for s in steps}

    # ------------------------------------------------------------ encoder
    def encode(self, values, scale: float, L: int) -> np.ndarray:
        z = np.zeros(self.slots, dtype=np.complex128)
        values = np.asarray(values, dtype=np.complex128).ravel()
        z[: values.size] = values
        v = np.zeros(self.n, dtype=np.complex128)
        v[self._i1] = z
        v[self._i2] = np.conj(z)
        coeff = np.rint((np.fft.fft(v) * np.conj(self._zeta) / self.n).real * scale)
        if np.max(np.abs(coeff)) >= 2**62:
            raise ValueError("encoded values are too large for encryption parameters")
        out = self._to_residues(coeff, L)
        self.ctx.ntt_forward_host(out, 0, L)
        return out

    def encode_many(self, rows, scale: float, L: int, special: bool = False) -> np.ndarray:
        """rows: [count][<=slots] -> [count][L][N]; one batched NTT on the GPU.  special=True appends a
        limb mod the special prime ([count][L+1][N], the plaintexts of HEGPU_MATVEC_DH: an encode at the
        key level restricted to the limbs q_0..q_{L-1}, P)."""
        rows = np.asarray(rows, dtype=np.complex128)
        cnt = rows.shape[0]
        v = np.zeros((cnt, self.n), dtype=np.complex128)
        z = np.zeros((cnt, self.slots), dtype=np.complex128)
        z[:, : rows.shape[1]] = rows
        v[:, self._i1] = z
        v[:, self._i2] = np.conj(z)
        coeff = np.rint((np.fft.fft(v, axis=1) * np.conj(self._zeta) / self.n).real * scale).astype(np.int64)
        out = np.empty((cnt, L, self.n), dtype=np.uint64)
        for i in range(L):
            out[:, i, :] = np.mod(coeff, np.int64(self.moduli[i])).astype(np.uint64)
        self.ctx.ntt_forward_host(out, 0, L)
        if special:
            sp = np.ascontiguousarray(np.mod(coeff, np.int64(self.moduli[self.K - 1])).astype(np.uint64))
            self.ctx.ntt_forward_host(sp, self.K - 1, 1)
            out = np.ascontiguousarray(np.concatenate([out, sp[:, None, :]], axis=1))
        return out

    def decode(self, plain_ntt: np.ndarray, scale: float) -> np.ndarray:
        L = plain_ntt.shape[0]
        c = np.ascontiguousarray(plain_ntt).copy()
        self.ctx.ntt_inverse_host(c, 0, L)
        mods = self.moduli[:L]
        Q = 1
        for q in mods:
            Q *= q
        acc = np.zeros(self.n, dtype=object)
        for i, q in enumerate(mods):
            Qi = Q // q
            acc = acc + c[i].astype(object) * (Qi * pow(Qi % q, q - 2, q))
        acc = acc % Q
        half = Q // 2
        m = np.array([float(x - Q) if x > half else float(x) for x in acc])
        return (np.fft.ifft(m * self._zeta) * self.n / scale)[self._i1]

    def decrypt_decode_many(self, cts: np.ndarray, scale: float) -> np.ndarray:
        """Decrypt and decode a batch of size-2 ciphertexts [B][2][L][N] -> complex slots [B][N/2], from limb 0
        alone: exact (before the embedding FFT) when every coefficient of the plaintext lies in (-q_0/2, q_0/2) --
        true for the bench workloads (|message| * scale < 2^50 against a 60-bit q_0); `decode` reconstructs over
        the whole basis.  One batched multiply and one batched INTT on the GPU."""
        ctx = self.ctx
        B, size, L, n = cts.shape
        assert size == 2
        A = ctx.upload_ct(np.ascontiguousarray(cts), 1.0, size_cap=2, L_cap=L)
        ctx.multiply_plain(A, A, ctx.upload_pt(np.ascontiguousarray(self.secret[:L]), 1.0, L_cap=L), 0)
        q0 = np.uint64(self.moduli[0])
        m = np.ascontiguousarray((cts[:, 0, 0] + A.download()[:, 1, 0]) % q0)  # c0 + c1*s mod q_0, NTT form
        ctx.ntt_inverse_host(m, 0, 1)
        v = m.astype(np.int64)
        v = np.where(v > np.int64(self.moduli[0] // 2), v - np.int64(self.moduli[0]), v).astype(np.float64)
        return (np.fft.ifft(v * self._zeta, axis=1) * self.n / scale)[:, self._i1]

    # ------------------------------------------------------------ encrypt / decrypt
    def encrypt_many(self, plains: np.ndarray) -> np.ndarray:
        """Encryptor::encrypt_symmetric: plains [B][L][N] (NTT form) -> [B][2][L][N]"""
        B, L, _ = plains.shape
        a = self._uniform(L, (B,))
        e = np.stack([self._noise_ntt(L) for _ in range(B)])
        c0 = self._neg_as_plus_e(a, e, L)
        ctx = self.ctx
        two = np.ascontiguousarray(np.stack([c0, a], axis=1))
        A = ctx.upload_ct(two, 1.0, size_cap=2, L_cap=L)
        ctx.add_plain(A, A, ctx.upload_pt(np.ascontiguousarray(plains), 1.0, L_cap=L), -1)
        return A.download()

    def decrypt(self, ct: np.ndarray) -> np.ndarray:
        """Decryptor: sum_k c_k s^k for one ciphertext [size][L][N] -> plaintext [L][N] (NTT form)"""
        ctx = self.ctx
        size, L, _ = ct.shape
        S = ctx.upload_pt(np.ascontiguousarray(self.secret[:L]), 1.0, L_cap=L)
        acc = np.ascontiguousarray(np.stack([ct[size - 1], ct[size - 1]])[None])
        A = ctx.upload_ct(acc, 1.0, size_cap=2, L_cap=L)
        for k in range(size - 2, -1, -1):
            ctx.multiply_plain(A, A, S, 0)
            A.scale = 1.0
            ctx.add_plain(A, A, ctx.upload_pt(np.ascontiguousarray(ct[k]), 1.0, L_cap=L), 0)
        return np.ascontiguousarray(A.download()[0, 0])
