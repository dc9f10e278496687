"""Reader / writer of the "HEGVEC1" vector container that tools/seal_golden.cpp emits from real SEAL 4.1
(test infrastructure; format documented at the top of that file), and the list of operations a vector file pins.

`run_pin(vec, ev)` replays every operation of the file on an evaluator adapter `ev` (the oracle or the GPU) with the
file's own keys and input ciphertexts and returns the names whose outputs differ -- the harness is the same for
vectors written by SEAL and for the self-test vectors `write_with_oracle` produces."""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"HEGVEC1\0"
DTYPES = {0: np.uint64, 1: np.float64, 2: np.int64}
CODES = {np.dtype(np.uint64): 0, np.dtype(np.float64): 1, np.dtype(np.int64): 2}


def read(path) -> dict:
    out = {}
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{path}: not a HEGVEC1 file")
        (cnt,) = struct.unpack("<I", f.read(4))
        for _ in range(cnt):
            (nl,) = struct.unpack("<I", f.read(4))
            name = f.read(nl).decode()
            dtype, nd = struct.unpack("<II", f.read(8))
            dims = struct.unpack(f"<{nd}Q", f.read(8 * nd))
            count = int(np.prod(dims)) if nd else 1
            out[name] = np.frombuffer(f.read(8 * count), dtype=DTYPES[dtype]).reshape(dims).copy()
    return out


def write(path, records: dict):
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(records)))
        for name, arr in records.items():
            arr = np.ascontiguousarray(arr)
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)) + nb)
            f.write(struct.pack("<II", CODES[arr.dtype], arr.ndim))
            f.write(struct.pack(f"<{arr.ndim}Q", *arr.shape))
            f.write(arr.tobytes())


def galois_keys(vec) -> dict:
    return {int(e): vec[f"galois_key.{int(e)}"] for e in vec["galois_elts"]}


# (record name, inputs, how the evaluator adapter computes it)
def operations(vec):
    ops = [
        ("add", lambda ev: ev.add(vec["ct_a"], vec["ct_b"])),
        ("sub", lambda ev: ev.sub(vec["ct_a"], vec["ct_b"])),
        ("negate", lambda ev: ev.negate(vec["ct_a"])),
        ("add_plain", lambda ev: ev.add_plain(vec["ct_a"], vec["pt"])),
        ("sub_plain", lambda ev: ev.sub_plain(vec["ct_a"], vec["pt"])),
        ("multiply_plain", lambda ev: ev.multiply_plain(vec["ct_a"], vec["pt"])),
        ("mod_switch", lambda ev: ev.mod_switch(vec["ct_a"])),
        ("multiply", lambda ev: ev.multiply(vec["ct_a"], vec["ct_b"])),
        ("square", lambda ev: ev.square(vec["ct_a"])),
        ("relinearize", lambda ev: ev.relinearize(vec["multiply"])),
        ("rescale", lambda ev: ev.rescale(vec["relinearize"])),
    ]
    for st in vec["rotate_steps"]:
        st = int(st)
        ops.append((f"rotate.{st}", lambda ev, st=st: ev.rotate(vec["ct_a"], st)))
        ops.append((f"rotate_low.{st}", lambda ev, st=st: ev.rotate(vec["rescale"], st)))
    return ops


def run_pin(vec, ev) -> list[str]:
    """Names of the operations whose output differs from the file's (empty = pinned)."""
    bad = []
    for name, fn in operations(vec):
        if not np.array_equal(fn(ev), vec[name]):
            bad.append(name)
    return bad


class OracleEvaluator:
    """Adapter: the CPU oracle driven with the file's keys."""

    def __init__(self, vec):
        from oracle import oracle as orc

        self.o = orc.Oracle(int(vec["n"][0]), [int(q) for q in vec["moduli"]])
        self.rk, self.gk = vec["relin_key"], galois_keys(vec)

    def add(self, a, b):
        return self.o.add(a, b)

    def sub(self, a, b):
        return self.o.sub(a, b)

    def negate(self, a):
        return self.o.negate(a)

    def add_plain(self, a, p):
        return self.o.add_plain(a, p)

    def sub_plain(self, a, p):
        return self.o.sub_plain(a, p)

    def multiply_plain(self, a, p):
        return self.o.multiply_plain(a, p)

    def mod_switch(self, a):
        return self.o.mod_switch(a)

    def multiply(self, a, b):
        return self.o.multiply(a, b)

    def square(self, a):
        return self.o.square(a)

    def relinearize(self, a):
        return self.o.relinearize(a, self.rk)

    def rescale(self, a):
        return self.o.rescale(a)

    def rotate(self, a, st):
        return a.copy() if st == 0 else self.o.rotate(a, st, self.gk)[0]


class GpuEvaluator:
    """Adapter: libhegpu.so through the C ABI (ctypes binding), batch of one ciphertext."""

    def __init__(self, vec, hg):
        self.vec = vec
        self.ctx = hg.Context(int(vec["n"][0]), [int(q) for q in vec["moduli"]])
        self.ctx.load_relin_key(vec["relin_key"])
        self.ctx.load_galois_keys(galois_keys(vec))
        self.scale = float(vec["ct_a.scale"][0])

    def _up(self, a, scale=None):
        return self.ctx.upload_ct(a[None], self.scale if scale is None else scale)

    def _bin(self, fn, a, b):
        out = self.ctx.ct(1)
        fn(out, self._up(a), self._up(b))
        return out.download()[0]

    def add(self, a, b):
        return self._bin(self.ctx.add, a, b)

    def sub(self, a, b):
        return self._bin(self.ctx.sub, a, b)

    def multiply(self, a, b):
        return self._bin(self.ctx.multiply, a, b)

    def _un(self, fn, a, scale=None):
        out = self.ctx.ct(1)
        fn(out, self._up(a, scale))
        return out.download()[0]

    def negate(self, a):
        return self._un(self.ctx.negate, a)

    def square(self, a):
        return self._un(self.ctx.square, a)

    def mod_switch(self, a):
        return self._un(self.ctx.mod_switch_to_next, a)

    def relinearize(self, a):
        return self._un(self.ctx.relinearize, a, self.scale * self.scale)

    def rescale(self, a):
        return self._un(self.ctx.rescale_to_next, a, self.scale * self.scale)

    def _pl(self, fn, a, p):
        out = self.ctx.ct(1)
        fn(out, self._up(a), self.ctx.upload_pt(p[None], self.scale), 0)
        return out.download()[0]

    def add_plain(self, a, p):
        return self._pl(self.ctx.add_plain, a, p)

    def sub_plain(self, a, p):
        return self._pl(self.ctx.sub_plain, a, p)

    def multiply_plain(self, a, p):
        return self._pl(self.ctx.multiply_plain, a, p)

    def rotate(self, a, st):
        out = self.ctx.ct(1)
        self.ctx.rotate_vector(out, self._up(a), st)
        return out.download()[0]


def write_with_oracle(path, n=4096, bits=(36, 36, 37), seed=3):
    """Self-test vectors in the SEAL tool's format, produced by the oracle (NOT a pin: it only proves that the
    harness consumes the format and replays every operation).  Default Galois key set, as the SEAL tool."""
    from oracle import oracle as orc

    moduli = orc.coeff_modulus_create(n, bits)
    o = orc.Oracle(n, moduli)
    K, L = len(moduli), len(moduli) - 1
    rng = np.random.default_rng(seed)
    s = o.sample_secret(seed)
    rec = {"moduli": np.array(moduli, dtype=np.uint64), "bits": np.array(bits, dtype=np.int64), "n": np.array([n], dtype=np.int64),
           "secret_key": s, "relin_key": o.gen_relin_key(seed + 1, s)}
    steps = [1, -1, 2, 4, -4, 8, 64, n // 4]
    gk = o.gen_galois_keys_for_steps(seed + 2, s, steps)
    for e, k in gk.items():
        rec[f"galois_key.{e}"] = k
    rec["galois_elts"] = np.array(sorted(gk), dtype=np.int64)
    scale = float(2**20)  # small enough for the product of two fresh ciphertexts on short chains

    def residues(*prefix):
        a = np.empty(prefix + (L, n), dtype=np.uint64)
        for i, q in enumerate(moduli[:L]):
            a[..., i, :] = rng.integers(0, q, size=prefix + (n,), dtype=np.uint64)
        return a

    pt = residues()
    a, b = o.encrypt_symmetric(seed + 3, s, residues()), o.encrypt_symmetric(seed + 4, s, residues())
    rec.update({"ct_a": a, "ct_a.scale": np.array([scale]), "ct_b": b, "ct_b.scale": np.array([scale]), "pt": pt,
                "pt.scale": np.array([scale])})
    vec = dict(rec)
    vec["rotate_steps"] = np.array([1, -1, 2, 64, 7, -5, n // 4, 0], dtype=np.int64)
    ev = OracleEvaluator(vec)
    for name, fn in operations(vec):
        vec[name] = fn(ev)
    write(path, vec)
    return path
