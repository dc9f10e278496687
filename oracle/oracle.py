"""ctypes binding of the CPU oracle (oracle/ckks_oracle.c).

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (real SEAL 4.1 is not available here).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module; the product package never does.

Arrays are numpy uint64 in SEAL's layout: ciphertext [size][L][N], plaintext [L][N],
key-switching key [Lmax][2][K][N].
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libckks_oracle.so")

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ckks_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libckks_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        L = _lib
        L.orc_ctx_create.restype = C.c_void_p
        L.orc_ctx_create.argtypes = [C.c_uint32, u64p, C.c_uint32]
        L.orc_ctx_free.argtypes = [C.c_void_p]
        L.orc_psi.restype = C.c_uint64
        L.orc_psi.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_get_primes.argtypes = [C.c_uint64, C.c_int, C.c_uint32, u64p]
        L.orc_coeff_modulus_create.argtypes = [C.c_uint32, C.POINTER(C.c_int), C.c_uint32, u64p]
        L.orc_is_prime.argtypes = [C.c_uint64]
        L.orc_galois_elt_from_step.restype = C.c_uint32
        L.orc_galois_elt_from_step.argtypes = [C.c_uint32, C.c_int]
        L.orc_naf.argtypes = [C.c_int, C.POINTER(C.c_int)]
        L.orc_max_threads.restype = C.c_int
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def coeff_modulus_create(n: int, bits) -> list[int]:
    bits = list(bits)
    arr = (C.c_int * len(bits))(*bits)
    out = np.zeros(len(bits), dtype=np.uint64)
    rc = lib().orc_coeff_modulus_create(n, arr, len(bits), _p(out))
    if rc:
        raise ValueError("failed to find enough qualifying primes")
    return [int(x) for x in out]


def get_primes(factor: int, bits: int, count: int) -> list[int]:
    out = np.zeros(count, dtype=np.uint64)
    if lib().orc_get_primes(factor, bits, count, _p(out)):
        raise ValueError("failed to find enough qualifying primes")
    return [int(x) for x in out]


def galois_elt_from_step(n: int, step: int) -> int:
    return int(lib().orc_galois_elt_from_step(n, step))


def naf(value: int) -> list[int]:
    buf = (C.c_int * 40)()
    cnt = lib().orc_naf(value, buf)
    return [buf[i] for i in range(cnt)]


class Oracle:
    """One CKKS parameter set: ring degree n and the full prime chain (last = special prime)."""

    def __init__(self, n: int, moduli):
        self.n = int(n)
        self.moduli = [int(q) for q in moduli]
        self.K = len(self.moduli)
        arr = np.array(self.moduli, dtype=np.uint64)
        self._h = lib().orc_ctx_create(self.n, _p(arr), self.K)
        if not self._h:
            raise ValueError("invalid parameters")
        self._h = C.c_void_p(self._h)

    def __del__(self):
        try:
            if self._h:
                lib().orc_ctx_free(self._h)
        except Exception:
            pass

    # ---- helpers
    def psi(self, i: int) -> int:
        return int(lib().orc_psi(self._h, i))

    def _out(self, *shape):
        return np.empty(shape, dtype=np.uint64)

    # ---- NTT
    def ntt_fwd(self, mod_index: int, a: np.ndarray) -> np.ndarray:
        r = np.ascontiguousarray(a, dtype=np.uint64).copy()
        lib().orc_ntt_fwd(self._h, C.c_uint32(mod_index), _p(r))
        return r

    def ntt_inv(self, mod_index: int, a: np.ndarray) -> np.ndarray:
        r = np.ascontiguousarray(a, dtype=np.uint64).copy()
        lib().orc_ntt_inv(self._h, C.c_uint32(mod_index), _p(r))
        return r

    # ---- element-wise
    def negate(self, a):
        a = np.ascontiguousarray(a)
        o = np.empty_like(a)
        lib().orc_negate(self._h, C.c_uint32(a.shape[1]), _p(a), C.c_uint32(a.shape[0]), _p(o))
        return o

    def _binary(self, fn, a, b):
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b)
        assert a.shape[1:] == b.shape[1:]
        o = self._out(max(a.shape[0], b.shape[0]), a.shape[1], self.n)
        fn(self._h, C.c_uint32(a.shape[1]), _p(a), C.c_uint32(a.shape[0]), _p(b), C.c_uint32(b.shape[0]), _p(o))
        return o

    def add(self, a, b):
        return self._binary(lib().orc_add, a, b)

    def sub(self, a, b):
        return self._binary(lib().orc_sub, a, b)

    def _plain(self, fn, ct, pt):
        ct = np.ascontiguousarray(ct)
        pt = np.ascontiguousarray(pt)
        assert pt.shape == ct.shape[1:]
        o = np.empty_like(ct)
        fn(self._h, C.c_uint32(ct.shape[1]), _p(ct), C.c_uint32(ct.shape[0]), _p(pt), _p(o))
        return o

    def add_plain(self, ct, pt):
        return self._plain(lib().orc_add_plain, ct, pt)

    def sub_plain(self, ct, pt):
        return self._plain(lib().orc_sub_plain, ct, pt)

    def multiply_plain(self, ct, pt):
        return self._plain(lib().orc_multiply_plain, ct, pt)

    def multiply(self, a, b):
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b)
        o = self._out(a.shape[0] + b.shape[0] - 1, a.shape[1], self.n)
        lib().orc_multiply(self._h, C.c_uint32(a.shape[1]), _p(a), C.c_uint32(a.shape[0]), _p(b), C.c_uint32(b.shape[0]), _p(o))
        return o

    def square(self, a):
        a = np.ascontiguousarray(a)
        o = self._out(2 * a.shape[0] - 1, a.shape[1], self.n)
        lib().orc_square(self._h, C.c_uint32(a.shape[1]), _p(a), C.c_uint32(a.shape[0]), _p(o))
        return o

    # ---- levels
    def rescale(self, ct):
        ct = np.ascontiguousarray(ct)
        o = self._out(ct.shape[0], ct.shape[1] - 1, self.n)
        lib().orc_rescale(self._h, C.c_uint32(ct.shape[1]), _p(ct), C.c_uint32(ct.shape[0]), _p(o))
        return o

    def mod_switch(self, ct):
        ct = np.ascontiguousarray(ct)
        o = self._out(ct.shape[0], ct.shape[1] - 1, self.n)
        lib().orc_mod_switch(self._h, C.c_uint32(ct.shape[1]), _p(ct), C.c_uint32(ct.shape[0]), _p(o))
        return o

    # ---- galois / key switch
    def galois_table(self, elt: int) -> np.ndarray:
        t = np.empty(self.n, dtype=np.uint32)
        lib().orc_galois_table(C.c_uint32(self.n), C.c_uint32(elt), t.ctypes.data_as(u32p))
        return t

    def apply_galois_ntt(self, polys, elt: int):
        polys = np.ascontiguousarray(polys)
        limbs = polys.size // self.n
        o = np.empty_like(polys)
        lib().orc_apply_galois_ntt(self._h, C.c_uint32(limbs), C.c_uint32(elt), _p(polys), _p(o))
        return o

    def switch_key(self, ct, target, key):
        r = np.ascontiguousarray(ct).copy()
        target = np.ascontiguousarray(target)
        lib().orc_switch_key(self._h, C.c_uint32(r.shape[1]), _p(r), _p(target), _p(key))
        return r

    def relinearize(self, ct3, rk):
        ct3 = np.ascontiguousarray(ct3)
        assert ct3.shape[0] == 3
        o = self._out(2, ct3.shape[1], self.n)
        lib().orc_relinearize(self._h, C.c_uint32(ct3.shape[1]), _p(ct3), _p(rk), _p(o))
        return o

    def apply_galois(self, ct, elt: int, key):
        ct = np.ascontiguousarray(ct)
        assert ct.shape[0] == 2
        o = np.empty_like(ct)
        lib().orc_apply_galois(self._h, C.c_uint32(ct.shape[1]), _p(ct), C.c_uint32(elt), _p(key), _p(o))
        return o

    def rotate(self, ct, steps: int, galois_keys: dict):
        """SEAL rotate_vector with NAF fallback. galois_keys: {elt: key array}.
        Returns (ct_out, number_of_key_switches). Raises on 'Galois key not present'."""
        ct = np.ascontiguousarray(ct)
        elts = np.array(list(galois_keys.keys()), dtype=np.uint32)
        keys = (u64p * len(elts))(*[_p(galois_keys[int(e)]) for e in elts])
        o = np.empty_like(ct)
        rc = lib().orc_rotate(self._h, C.c_uint32(ct.shape[1]), _p(ct), C.c_int(steps), C.c_uint32(len(elts)),
                              elts.ctypes.data_as(u32p), keys, _p(o))
        if rc < 0:
            raise ValueError("Galois key not present")
        return o, rc

    # ---- client side
    def sample_secret(self, seed: int):
        s = self._out(self.K, self.n)
        lib().orc_sample_secret(self._h, C.c_uint64(seed), _p(s))
        return s

    def gen_relin_key(self, seed: int, s):
        k = self._out(self.K - 1, 2, self.K, self.n)
        lib().orc_gen_relin_key(self._h, C.c_uint64(seed), _p(s), _p(k))
        return k

    def gen_galois_key(self, seed: int, s, elt: int):
        k = self._out(self.K - 1, 2, self.K, self.n)
        lib().orc_gen_galois_key(self._h, C.c_uint64(seed), _p(s), C.c_uint32(elt), _p(k))
        return k

    def gen_galois_keys_for_steps(self, seed: int, s, steps):
        return {galois_elt_from_step(self.n, st): self.gen_galois_key(seed + 17 * i + 1, s, galois_elt_from_step(self.n, st))
                for i, st in enumerate(steps)}

    def default_galois_elts(self):
        """SEAL create_galois_keys() with no list: 3^(2^i), 3^(-2^i) for i < log2(N)-1.
        (The column-swap element 2N-1 is not used by rotate_vector and is omitted.)"""
        logn = self.n.bit_length() - 1
        steps = []
        for i in range(logn - 1):
            steps += [1 << i, -(1 << i)]
        return steps

    def encrypt_symmetric(self, seed: int, s, plain):
        plain = np.ascontiguousarray(plain)
        L = plain.shape[0]
        ct = self._out(2, L, self.n)
        lib().orc_encrypt_symmetric(self._h, C.c_uint64(seed), C.c_uint32(L), _p(s), _p(plain), _p(ct))
        return ct

    def decrypt(self, ct, s):
        ct = np.ascontiguousarray(ct)
        o = self._out(ct.shape[1], self.n)
        lib().orc_decrypt(self._h, C.c_uint32(ct.shape[1]), _p(ct), C.c_uint32(ct.shape[0]), _p(s), _p(o))
        return o

    # ---- composites
    def matvec_bsgs(self, cts, n1: int, n2: int, pts, baby_keys, giant_keys, threads: int = 1, fast: bool = False,
                    hoist: bool | None = None, lazy: bool | None = None, rescale: bool = True, g_first: int = 0,
                    dh: bool = False):
        """cts [B][2][L][N]; pts [n1*n2][L][N]; baby_keys[k], giant_keys[g'] lists (unused entries may be None).
        fast=True restates HEGPU_MATVEC_HOIST|LAZY; hoist / lazy select them separately; g_first is the first
        global giant step of a diagonal-sharded call.  dh=True restates HEGPU_MATVEC_DH (double-hoisted); pts
        then carries a limb mod the special prime: [n1*n2][L+1][N]."""
        cts = np.ascontiguousarray(cts)
        pts = np.ascontiguousarray(pts)
        B, _, L, _ = cts.shape
        assert pts.shape[1] == (L + 1 if dh else L)
        hoist = fast if hoist is None else hoist
        lazy = fast if lazy is None else lazy
        null = C.cast(None, u64p)
        bk = (u64p * n1)(*[null if k is None else _p(k) for k in baby_keys])
        gk = (u64p * n2)(*[null if k is None else _p(k) for k in giant_keys])
        o = self._out(B, 2, L - 1 if rescale else L, self.n)
        if dh:
            lib().orc_matvec_bsgs_dh(self._h, C.c_uint32(L), C.c_uint32(B), _p(cts), C.c_uint32(n1), C.c_uint32(n2),
                                     C.c_uint32(g_first), _p(pts), bk, gk, C.c_int(4 if rescale else 0), _p(o), C.c_int(threads))
        elif not hoist and not lazy and rescale and g_first == 0:
            lib().orc_matvec_bsgs(self._h, C.c_uint32(L), C.c_uint32(B), _p(cts), C.c_uint32(n1), C.c_uint32(n2), _p(pts), bk, gk,
                                  _p(o), C.c_int(threads))
        else:
            flags = (1 if hoist else 0) | (2 if lazy else 0) | (4 if rescale else 0)
            lib().orc_matvec_bsgs_ex(self._h, C.c_uint32(L), C.c_uint32(B), _p(cts), C.c_uint32(n1), C.c_uint32(n2),
                                     C.c_uint32(g_first), _p(pts), bk, gk, C.c_int(flags), _p(o), C.c_int(threads))
        return o


def max_threads() -> int:
    return int(lib().orc_max_threads())
