"""Shared fixtures: oracle context + keys + ciphertext builders (test infrastructure)."""
from __future__ import annotations

import functools

import numpy as np

import ckks_ref as ref
from oracle import oracle as orc


def rand_residues(rng, moduli, prefix, n):
    out = np.empty(tuple(prefix) + (len(moduli), n), dtype=np.uint64)
    for i, q in enumerate(moduli):
        out[..., i, :] = rng.integers(0, q, size=tuple(prefix) + (n,), dtype=np.uint64)
    return out


def ckks_tol(n_terms, N, scale, mode="exact", out_scale=None):
    """Decryption tolerance of a diagonal matvec with n_terms diagonals (DESIGN.md "Tolerance"), calibrated on
    the oracle (tools/calibrate_tolerance.py; VERDICT r1 asked for <= ~10x the measured error):

    mode "exact"   -- chain of SEAL primitives (rotate_vector per step, NAF chains).  SEAL's RNS digits are
                      non-centred ([0,q_j)), so the key-switch noise sum_j c_j*e_j/P carries a random-walk
                      component of amplitude ~sigma*sqrt(N)/2 whose energy lands in the slots whose roots lie next
                      to 1 -- the first slots, where the result lives: n_terms * sigma * N^1.5 / (8 * scale),
                      6x - 19x the measured error (4.5e-7 / 4.0e-6 / 5.1e-6 at (N, dim) = (8192, 16) /
                      (16384, 32) / (16384, 128)).
    mode "hoisted" -- HOIST, HOIST|LAZY and DH: the rotations act on the lifted digits, every key-switched term is
                      multiplied by a diagonal before it is summed, and the error behaves like independent noise:
                      2400 * sqrt(n_terms * N) / scale, 8x - 14x the measured error (1.0e-7 / 1.1e-7 / 2.3e-7 /
                      1.1e-6 at (8192, 16) / (16384, 32) / (16384, 128) / (32768, 512)).
    out_scale      -- scale of the rescaled result, when it differs from `scale` (chains whose data primes are
                      not close to the scale)."""
    # the final rescale rounds every coefficient: ~N / (4 * out_scale) per slot (secret of Hamming weight 2N/3);
    # only visible when the dropped prime is much larger than the scale (3.4e-6 measured at N = 8192, out_scale 2^30)
    round_term = 4.0 * N / out_scale if out_scale else 0.0
    if mode == "hoisted":
        return 2400.0 * (n_terms * N) ** 0.5 / scale + round_term
    return n_terms * 3.2 * N**1.5 / (8.0 * scale) + round_term


class Setup:
    """One parameter set with an oracle, an encoder, a secret key and lazily generated keys."""

    def __init__(self, n, bits, seed=1):
        self.n = n
        self.bits = list(bits)
        self.moduli = orc.coeff_modulus_create(n, self.bits)
        self.K = len(self.moduli)
        self.Lmax = self.K - 1
        self.o = orc.Oracle(n, self.moduli)
        self.enc = ref.Encoder(n, self.moduli, self.o.ntt_fwd, self.o.ntt_inv)
        self.seed = seed
        self.s = self.o.sample_secret(seed)
        self._rk = None
        self._gk = {}

    @property
    def rk(self):
        if self._rk is None:
            self._rk = self.o.gen_relin_key(self.seed + 1, self.s)
        return self._rk

    def gk(self, steps):
        """{elt: key} for the given rotation steps."""
        out = {}
        for st in steps:
            elt = orc.galois_elt_from_step(self.n, st)
            if elt not in self._gk:
                self._gk[elt] = self.o.gen_galois_key(self.seed + 1000 + elt, self.s, elt)
            out[elt] = self._gk[elt]
        return out

    def encrypt(self, values, scale, L=None, seed=0):
        L = self.Lmax if L is None else L
        return self.o.encrypt_symmetric(self.seed + 5000 + seed, self.s, self.enc.encode(values, scale, L))

    def decrypt(self, ct, scale):
        return self.enc.decode(self.o.decrypt(ct, self.s), scale)


@functools.lru_cache(maxsize=None)
def setup(n, bits, seed=1):
    return Setup(n, bits, seed)
