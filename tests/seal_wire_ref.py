"""Independent Python restatement of the pieces of SEAL 4.1's wire format that host/seal_wire.cpp implements (test
infrastructure): BLAKE2b with a parameter block (RFC 7693), BLAKE2xb, the Blake2xb PRNG, sample_poly_uniform, the
16-byte container and the member layouts of EncryptionParameters / Ciphertext / KSwitchKeys.  Written from the same
published SEAL sources as the C++ (format fidelity vs real SEAL is unpinned -- no SEAL here), but sharing no code with
it: the two must agree byte for byte, and BLAKE2b itself is pinned by hashlib."""
from __future__ import annotations

import ctypes
import struct
import zlib

import numpy as np

M64 = (1 << 64) - 1
IV = [0x6A09E667F3BCC908, 0xBB67AE8584CAA73B, 0x3C6EF372FE94F82B, 0xA54FF53A5F1D36F1,
      0x510E527FADE682D1, 0x9B05688C2B3E6C1F, 0x1F83D9ABFB41BD6B, 0x5BE0CD19137E2179]
SIGMA = [
    [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15], [14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3],
    [11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4], [7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8],
    [9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13], [2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9],
    [12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11], [13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10],
    [6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5], [10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0],
]


def _rotr(x, r):
    return ((x >> r) | (x << (64 - r))) & M64


def _compress(h, block, t, last):
    m = struct.unpack("<16Q", block)
    v = h + IV[:]
    v[12] ^= t & M64
    v[13] ^= t >> 64
    if last:
        v[14] ^= M64
    for r in range(12):
        s = SIGMA[r % 10]
        for i, (a, b, c, d) in enumerate(((0, 4, 8, 12), (1, 5, 9, 13), (2, 6, 10, 14), (3, 7, 11, 15), (0, 5, 10, 15), (1, 6, 11, 12),
                                          (2, 7, 8, 13), (3, 4, 9, 14))):
            v[a] = (v[a] + v[b] + m[s[2 * i]]) & M64
            v[d] = _rotr(v[d] ^ v[a], 32)
            v[c] = (v[c] + v[d]) & M64
            v[b] = _rotr(v[b] ^ v[c], 24)
            v[a] = (v[a] + v[b] + m[s[2 * i + 1]]) & M64
            v[d] = _rotr(v[d] ^ v[a], 16)
            v[c] = (v[c] + v[d]) & M64
            v[b] = _rotr(v[b] ^ v[c], 63)
    return [h[i] ^ v[i] ^ v[i + 8] for i in range(8)]


def blake2b_param(param: bytes, data: bytes, key: bytes = b"") -> bytes:
    """BLAKE2b with an explicit 64-byte parameter block (digest length = param[0], key length = param[1])."""
    assert len(param) == 64 and param[1] == len(key)
    h = [IV[i] ^ struct.unpack_from("<Q", param, 8 * i)[0] for i in range(8)]
    msg = (key.ljust(128, b"\0") if key else b"") + data
    blocks = [msg[i:i + 128] for i in range(0, len(msg), 128)] or [b""]
    t = 0
    for blk in blocks[:-1]:
        t += 128
        h = _compress(h, blk, t, False)
    t += len(blocks[-1])
    h = _compress(h, blocks[-1].ljust(128, b"\0"), t, True)
    return struct.pack("<8Q", *h)[:param[0]]


def param_block(digest, keylen=0, fanout=1, depth=1, leaf=0, node_offset=0, xof=0, node_depth=0, inner=0) -> bytes:
    return struct.pack("<BBBBIIIBB", digest, keylen, fanout, depth, leaf, node_offset, xof, node_depth, inner).ljust(64, b"\0")


def blake2xb(outlen: int, data: bytes, key: bytes) -> bytes:
    root = blake2b_param(param_block(64, len(key), xof=outlen), data, key)
    out, i = b"", 0
    while len(out) < outlen:
        blk = min(64, outlen - len(out))
        out += blake2b_param(param_block(blk, 0, 0, 0, 64, i, outlen, 0, 64), root)
        i += 1
    return out


class Prng:
    """seal::Blake2xbPRNG: 4096-byte buffers, buffer k = blake2xb(4096, counter k as 8 LE bytes, key = 64-byte seed)."""

    def __init__(self, seed_words):
        self.seed = struct.pack("<8Q", *[int(x) for x in seed_words])
        self.counter, self.buf, self.pos = 0, b"", 0

    def generate(self, count: int) -> bytes:
        out = b""
        while len(out) < count:
            if self.pos == len(self.buf):
                self.buf = blake2xb(4096, struct.pack("<Q", self.counter), self.seed)
                self.counter += 1
                self.pos = 0
            take = min(count - len(out), len(self.buf) - self.pos)
            out += self.buf[self.pos:self.pos + take]
            self.pos += take
        return out


def sample_poly_uniform(prng: Prng, moduli, n: int) -> np.ndarray:
    words = np.frombuffer(prng.generate(len(moduli) * n * 8), dtype="<u8").reshape(len(moduli), n).copy()
    for j, q in enumerate(moduli):
        q = int(q)
        max_multiple = M64 - (M64 % q) - 1
        for i in np.nonzero(words[j] >= np.uint64(max_multiple))[0]:
            r = int(words[j, i])
            while r >= max_multiple:
                r = struct.unpack("<Q", prng.generate(8))[0]
            words[j, i] = r
        words[j] %= np.uint64(q)
    return words


# ------------------------------------------------------------------ container and members
def _zstd():
    try:
        return ctypes.CDLL("libzstd.so.1")
    except OSError:
        return None


def zstd_compress(data: bytes) -> bytes:
    z = _zstd()
    z.ZSTD_compressBound.restype = ctypes.c_size_t
    z.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
    z.ZSTD_compress.restype = ctypes.c_size_t
    z.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]
    cap = z.ZSTD_compressBound(len(data))
    buf = ctypes.create_string_buffer(cap)
    n = z.ZSTD_compress(buf, cap, data, len(data), 3)
    return buf.raw[:n]


def zstd_decompress(data: bytes, cap: int) -> bytes:
    z = _zstd()
    z.ZSTD_decompress.restype = ctypes.c_size_t
    z.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
    buf = ctypes.create_string_buffer(cap)
    n = z.ZSTD_decompress(buf, cap, data, len(data))
    return buf.raw[:n]


def wrap(members: bytes, mode: int = 0) -> bytes:
    body = members if mode == 0 else zlib.compress(members) if mode == 1 else zstd_compress(members)
    return struct.pack("<HBBBBHQ", 0xA15E, 0x10, 4, 1, mode, 0, 16 + len(body)) + body


def unwrap(buf: bytes, off: int = 0):
    magic, hs, vmaj, vmin, mode, _, size = struct.unpack_from("<HBBBBHQ", buf, off)
    assert magic == 0xA15E and hs == 0x10
    body = buf[off + 16:off + size]
    if mode == 1:
        body = zlib.decompress(body)
    elif mode == 2:
        body = zstd_decompress(body, 1 << 28)
    return body, off + size


def parms_id(n, moduli, limbs, scheme=2, plain_modulus=0) -> bytes:
    import hashlib

    words = [scheme, n] + [int(q) for q in moduli[:limbs]] + [plain_modulus]
    return hashlib.blake2b(struct.pack(f"<{len(words)}Q", *words), digest_size=32).digest()


def save_parms(n, moduli, mode=0, scheme=2) -> bytes:
    m = struct.pack("<BQQ", scheme, n, len(moduli))
    for q in list(moduli) + [0]:
        m += wrap(struct.pack("<Q", int(q)))
    return wrap(m, mode)


def load_parms(buf, off=0):
    m, end = unwrap(buf, off)
    scheme, n, k = struct.unpack_from("<BQQ", m, 0)
    pos, vals = 17, []
    for _ in range(k + 1):
        mm, pos = unwrap(m, pos)
        vals.append(struct.unpack("<Q", mm)[0])
    return dict(scheme=scheme, n=n, moduli=vals[:-1], plain_modulus=vals[-1]), end


def ct_members(n, moduli, data: np.ndarray, scale: float, seed=None) -> bytes:
    size, limbs, _ = data.shape
    m = parms_id(n, moduli, limbs) + struct.pack("<BQQQQd", 1, size, n, limbs, 1, scale)
    if seed is None:
        m += wrap(struct.pack("<Q", data.size) + np.ascontiguousarray(data).astype("<u8").tobytes())
    else:
        assert size == 2
        half = np.ascontiguousarray(data[0]).astype("<u8")
        m += wrap(struct.pack("<Q", half.size) + half.tobytes())
        m += wrap(struct.pack("<B8Q", 1, *[int(x) for x in seed]))
    return m


def save_ciphertext(n, moduli, data, scale, seed=None, mode=0) -> bytes:
    return wrap(ct_members(n, moduli, data, scale, seed), mode)


def load_ciphertext(n, moduli, buf, off=0):
    m, end = unwrap(buf, off)
    pid = m[:32]
    ntt, size, nn, limbs, corr, scale = struct.unpack_from("<BQQQQd", m, 32)
    assert nn == n and pid == parms_id(n, moduli, limbs)
    arr, pos = unwrap(m, 32 + 41)
    (count,) = struct.unpack_from("<Q", arr, 0)
    words = np.frombuffer(arr, dtype="<u8", count=count, offset=8)
    total = size * limbs * n
    if count == total:
        data = words.reshape(size, limbs, n).copy()
    else:
        assert size == 2 and count == total // 2
        info, pos = unwrap(m, pos)
        assert info[0] == 1
        c1 = sample_poly_uniform(Prng(struct.unpack("<8Q", info[1:])), moduli[:limbs], n)
        data = np.stack([words.reshape(limbs, n), c1])
    return dict(data=data, scale=scale, size=size, limbs=limbs, ntt=bool(ntt), seeded=count != total), end


def save_kswitch_keys(n, moduli, key: np.ndarray, seeds=None, mode=0) -> bytes:
    """key [digits][2][K][n] (one index: RelinKeys); seeds [digits][8] or None."""
    K = len(moduli)
    m = parms_id(n, moduli, K) + struct.pack("<QQ", 1, key.shape[0])
    for j in range(key.shape[0]):
        m += wrap(wrap(ct_members(n, moduli, key[j], 1.0, None if seeds is None else seeds[j])))
    return wrap(m, mode)
