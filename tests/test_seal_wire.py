"""SEAL 4.1's wire format (host/seal_wire.cpp, libseal_wire.so) and the reference server's modes
served from the GPU evaluator (host/he_server.cpp).

CPU tier: BLAKE2b against hashlib (RFC 7693), BLAKE2xb / the PRNG / sample_poly_uniform / parms_id / the container and
the member layouts against the independent Python restatement tests/seal_wire_ref.py, in both directions and with every
compression mode.  GPU tier: this file plays src/demos/client.cpp -- parameters, relinearisation keys and SEEDED
symmetric ciphertexts serialized as SEAL's client sends them -- and checks the decrypted replies of server_side_simple,
server_side_batch_matmul and server_side_fft.  Format fidelity against real SEAL is unpinned (no SEAL in the image)."""
import ctypes as C
import hashlib
import os
import struct
import subprocess

import numpy as np
import pytest

import hegpu_loader
import seal_wire_ref as ref

LIB = os.path.join(hegpu_loader.PKG_DIR, "libseal_wire.so")
u8p, u64p, szp = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_size_t)


@pytest.fixture(scope="module")
def W():
    assert os.path.exists(LIB), "wire layer not built: run __graft_entry__.build()"
    w = C.CDLL(LIB)
    w.hewire_last_error.restype = C.c_char_p
    return w


def ck(W, rc):
    assert rc == 0, W.hewire_last_error().decode()


def taken(W, ptr, n, dtype=np.uint8):
    out = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n,)).copy()
    W.hewire_free(ptr)
    return out.view(dtype)


def b2(W, outlen, data, key=b""):
    out = C.create_string_buffer(outlen)
    ck(W, W.hewire_blake2b(out, C.c_size_t(outlen), data, C.c_size_t(len(data)), key, C.c_size_t(len(key))))
    return out.raw


def arr64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a, a.ctypes.data_as(u64p)


def test_blake2b_matches_hashlib(W):
    rng = np.random.default_rng(1)
    for n in (0, 1, 3, 64, 127, 128, 129, 255, 256, 257, 1000):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        for outlen in (1, 32, 48, 64):
            assert b2(W, outlen, data) == hashlib.blake2b(data, digest_size=outlen).digest()
        for klen in (1, 32, 64):
            key = rng.integers(0, 256, klen, dtype=np.uint8).tobytes()
            assert b2(W, 64, data, key) == hashlib.blake2b(data, digest_size=64, key=key).digest()
    assert b2(W, 64, b"abc").hex().startswith("ba80a53f981c4d0d6a2797b69f12f6e9")  # RFC 7693 appendix A
    # the Python restatement agrees with hashlib on parameter blocks hashlib can express (tree parameters, XOF length)
    for kw in (dict(), dict(leaf_size=64, inner_size=64, node_offset=5, fanout=0), dict(node_offset=(4096 << 32))):
        pb = ref.param_block(64, 0, kw.get("fanout", 1), 1, kw.get("leaf_size", 0), kw.get("node_offset", 0) & 0xFFFFFFFF,
                             kw.get("node_offset", 0) >> 32, 0, kw.get("inner_size", 0))
        assert ref.blake2b_param(pb, b"hello") == hashlib.blake2b(b"hello", digest_size=64, **kw).digest()


def test_blake2xb_prng_and_sampler_match_restatement(W):
    key = bytes(range(64))
    for outlen in (1, 64, 65, 200, 4096):
        out = C.create_string_buffer(outlen)
        ck(W, W.hewire_blake2xb(out, C.c_size_t(outlen), b"\x07\0\0\0\0\0\0\0", C.c_size_t(8), key, C.c_size_t(64)))
        assert out.raw == ref.blake2xb(outlen, b"\x07\0\0\0\0\0\0\0", key)
    seed, seedp = arr64(np.array([(i * 0x0123456789ABCDEF + 99) % (1 << 64) for i in range(8)], dtype=np.uint64))
    buf = C.create_string_buffer(10000)
    ck(W, W.hewire_prng_generate(seedp, C.c_size_t(10000), C.cast(buf, u8p)))
    p = ref.Prng(seed)
    assert buf.raw == p.generate(3) + p.generate(4093) + p.generate(5904)  # crosses two refills, any chunking
    # sampler: a tiny modulus makes rejections (words >= the largest multiple below 2^64) likely enough to be exercised
    moduli = [0xFFFFFFFFFFC0001, 0xFFFF8A0001, (1 << 63) + 29]
    mods, modp = arr64(moduli)
    n = 1024
    got = np.empty((3, n), dtype=np.uint64)
    ck(W, W.hewire_sample_poly_uniform(seedp, modp, C.c_size_t(3), C.c_size_t(n), got.ctypes.data_as(u64p)))
    want = ref.sample_poly_uniform(ref.Prng(seed), moduli, n)
    assert np.array_equal(got, want)
    assert all(int(got[j].max()) < moduli[j] for j in range(3))


def test_parms_roundtrip_and_parms_id(W):
    n, moduli = 8192, [0xFFFFFFFFFFE8001, 0xFFFFF4C001, 0xFFFFFDC001, 0xFFFFFFFFFFFC001]  # SURVEY 9.1 chain
    mods, modp = arr64(moduli)
    for limbs in (4, 3, 1):
        pid = (C.c_uint64 * 4)()
        ck(W, W.hewire_parms_id(C.c_uint64(n), modp, C.c_size_t(4), C.c_size_t(limbs), pid))
        assert bytes(pid) == ref.parms_id(n, moduli, limbs)
    modes = [0, 1] + ([2] if W.hewire_zstd_available() else [])
    for mode in modes:
        # Python writes, C++ reads
        blob = ref.save_parms(n, moduli, mode)
        used, scheme, nn, k, pm = C.c_size_t(), C.c_uint8(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        out = (C.c_uint64 * 64)()
        ck(W, W.hewire_load_parms(blob, C.c_size_t(len(blob)), C.byref(used), C.byref(scheme), C.byref(nn), C.byref(k), out, C.byref(pm)))
        assert (used.value, scheme.value, nn.value, k.value, pm.value) == (len(blob), 2, n, 4, 0) and list(out[:4]) == moduli
        # C++ writes, Python reads
        p, ln = u8p(), C.c_size_t()
        ck(W, W.hewire_save_parms(C.c_uint64(n), modp, C.c_size_t(4), mode, C.byref(p), C.byref(ln)))
        back, end = ref.load_parms(taken(W, p, ln.value).tobytes())
        assert end == ln.value and back == dict(scheme=2, n=n, moduli=moduli, plain_modulus=0)
    if 0 in modes:  # uncompressed streams are byte-identical in both implementations
        p, ln = u8p(), C.c_size_t()
        ck(W, W.hewire_save_parms(C.c_uint64(n), modp, C.c_size_t(4), 0, C.byref(p), C.byref(ln)))
        assert taken(W, p, ln.value).tobytes() == ref.save_parms(n, moduli, 0)
    # a corrupted header is refused with SEAL's message
    bad = bytearray(ref.save_parms(n, moduli, 0))
    bad[0] ^= 1
    used = C.c_size_t()
    assert W.hewire_load_parms(bytes(bad), C.c_size_t(len(bad)), C.byref(used), C.byref(C.c_uint8()), C.byref(C.c_uint64()), C.byref(C.c_uint64()),
                               (C.c_uint64 * 64)(), C.byref(C.c_uint64())) == 1
    assert b"SEALHeader" in W.hewire_last_error()


def _load_ct(W, n, moduli, blob):
    mods, modp = arr64(moduli)
    used, size, limbs, scale, seeded, data = C.c_size_t(), C.c_uint64(), C.c_uint64(), C.c_double(), C.c_int(), u64p()
    ck(W, W.hewire_load_ciphertext(C.c_uint64(n), modp, C.c_size_t(len(moduli)), blob, C.c_size_t(len(blob)), C.byref(used), C.byref(size), C.byref(limbs),
                                   C.byref(scale), C.byref(seeded), C.byref(data)))
    arr = taken(W, data, size.value * limbs.value * n * 8, np.uint64).reshape(size.value, limbs.value, n)
    return arr, scale.value, bool(seeded.value), used.value


def test_ciphertext_seeded_and_plain_both_directions(W):
    n, moduli = 4096, [0xFFFFFFFFFFC0001 % (1 << 36) | 1, 68719230977, 137438822401]  # any odd moduli do for the container
    moduli = [int(q) for q in moduli]
    rng = np.random.default_rng(3)
    L = 2
    seed = rng.integers(0, 1 << 63, 8, dtype=np.uint64)
    c1 = ref.sample_poly_uniform(ref.Prng(seed), moduli[:L], n)
    c0 = np.stack([rng.integers(0, moduli[i], n, dtype=np.uint64) for i in range(L)])
    data = np.stack([c0, c1])
    modes = [0, 1] + ([2] if W.hewire_zstd_available() else [])
    for mode in modes:
        for sd in (None, seed):
            blob = ref.save_ciphertext(n, moduli, data, 2.0**30, sd, mode)
            arr, scale, seeded, used = _load_ct(W, n, moduli, blob + b"trailing bytes of the next object")
            assert used == len(blob) and scale == 2.0**30 and seeded == (sd is not None)
            assert np.array_equal(arr, data)  # the C++ expansion of the seed equals the Python one
    # C++ writes (plain and seeded), Python reads
    mods, modp = arr64(moduli)
    d, dp = arr64(data)
    s, sp = arr64(seed)
    for sd in (None, sp):
        p, ln = u8p(), C.c_size_t()
        ck(W, W.hewire_save_ciphertext(C.c_uint64(n), modp, C.c_size_t(3), C.c_uint64(2), C.c_uint64(L), C.c_double(2.0**30), dp, sd, 0, C.byref(p), C.byref(ln)))
        blob = taken(W, p, ln.value).tobytes()
        assert blob == ref.save_ciphertext(n, moduli, data, 2.0**30, None if sd is None else seed, 0)
        back, end = ref.load_ciphertext(n, moduli, blob)
        assert end == len(blob) and np.array_equal(back["data"], data) and back["seeded"] == (sd is not None)
    # a ciphertext of another parameter set is refused
    other = [moduli[0], moduli[2], moduli[1]]
    with pytest.raises(AssertionError, match="ciphertext data is invalid"):
        _load_ct(W, n, other, ref.save_ciphertext(n, moduli, data, 1.0))


def test_kswitch_keys_seeded_roundtrip(W):
    n, moduli = 4096, [68718428161, 68719230977, 137438822401]
    K = 3
    rng = np.random.default_rng(4)
    seeds = rng.integers(0, 1 << 63, (K - 1, 8), dtype=np.uint64)
    key = np.empty((K - 1, 2, K, n), dtype=np.uint64)
    for j in range(K - 1):
        key[j, 0] = np.stack([rng.integers(0, moduli[i], n, dtype=np.uint64) for i in range(K)])
        key[j, 1] = ref.sample_poly_uniform(ref.Prng(seeds[j]), moduli, n)
    mods, modp = arr64(moduli)
    for sd in (None, seeds):
        blob = ref.save_kswitch_keys(n, moduli, key, sd, mode=1)
        used, idx, dig, flat = C.c_size_t(), C.c_uint64(), C.c_uint64(), u64p()
        ck(W, W.hewire_load_kswitch_keys(C.c_uint64(n), modp, C.c_size_t(K), blob, C.c_size_t(len(blob)), C.byref(used), C.byref(idx), C.byref(dig), C.byref(flat)))
        assert (used.value, idx.value, dig.value) == (len(blob), 1, K - 1)
        assert np.array_equal(taken(W, flat, key.size * 8, np.uint64).reshape(key.shape), key)
        k, kp = arr64(key)
        p, ln = u8p(), C.c_size_t()
        sdp = None if sd is None else arr64(sd)[1]
        ck(W, W.hewire_save_kswitch_keys(C.c_uint64(n), modp, C.c_size_t(K), kp, C.c_size_t(K - 1), sdp, 0, C.byref(p), C.byref(ln)))
        assert taken(W, p, ln.value).tobytes() == ref.save_kswitch_keys(n, moduli, key, sd, mode=0)


# ------------------------------------------------------------------ GPU tier: the reference client's requests
def _seeded_encrypt(S, values, scale, L, rng):
    """Encryptor::encrypt_symmetric(...).save(): c1 = expansion of a fresh seed (taken as NTT form), c0 = -(c1 s + e) + m."""
    seed = rng.integers(0, 1 << 63, 8, dtype=np.uint64)
    mods = S.moduli[:L]
    c1 = ref.sample_poly_uniform(ref.Prng(seed), mods, S.n)
    m = S.enc.encode(values, scale, L)
    e = rng.integers(-6, 7, S.n)
    c0 = np.empty_like(c1)
    for i, q in enumerate(mods):
        e_ntt = S.o.ntt_fwd(i, np.mod(e, q).astype(np.uint64)).astype(object)
        c0[i] = ((m[i].astype(object) - c1[i].astype(object) * S.s[i].astype(object) - e_ntt) % q).astype(np.uint64)
    return np.stack([c0, c1]), seed


def _serve(tmp_path, mode, request: bytes, args=()):
    exe = os.path.join(hegpu_loader.PKG_DIR, "he_host_test")
    req, rep = str(tmp_path / "request.bin"), str(tmp_path / "reply.bin")
    open(req, "wb").write(request)
    r = subprocess.run([exe, req, mode, rep] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    return open(rep, "rb").read()


def _replies(S, buf):
    out, off = [], 0
    while off < len(buf):
        c, off = ref.load_ciphertext(S.n, S.moduli, buf, off)
        assert not c["seeded"]
        out.append(c)
    return out


@pytest.mark.gpu
def test_server_side_simple(tmp_path):
    """client_side_simple (client.cpp:66-160): N = 8192 {60,40,40,60}, two complex operands, zstd-compressed request."""
    from fixtures import setup

    S = setup(8192, (60, 40, 40, 60))
    rng = np.random.default_rng(10)
    scale, L = 2.0**40, 3
    op1, op2 = complex(2.1, -9.5), complex(-5.3, -8.7)
    mode = 2 if ref._zstd() is not None else 1  # SEAL's default compr_mode is zstd
    req = ref.save_parms(S.n, S.moduli, mode) + ref.save_kswitch_keys(S.n, S.moduli, S.rk, None, mode)
    for v in (op1, op2):
        ct, seed = _seeded_encrypt(S, np.full(S.n // 2, v), scale, L, rng)
        req += ref.save_ciphertext(S.n, S.moduli, ct, scale, seed, mode)
    (res,) = _replies(S, _serve(tmp_path, "serve_simple", req))
    assert res["limbs"] == L - 1 and res["size"] == 2
    got = S.decrypt(res["data"], res["scale"])[0]
    assert abs(got - op1 * op2) < 1e-4, got


@pytest.mark.gpu
def test_server_side_batch_matmul_and_fft(tmp_path):
    from fixtures import setup

    S = setup(8192, (60, 40, 40, 60))
    rng = np.random.default_rng(11)
    scale, L = 2.0**40, 3
    A, B = rng.uniform(-2, 2, (2, 3)), rng.uniform(-2, 2, (3, 2))
    req = ref.save_parms(S.n, S.moduli, 1) + ref.save_kswitch_keys(S.n, S.moduli, S.rk, None, 1)
    for M in (A, B):  # column-major element order (he_linalg.h Matrix)
        for j in range(M.shape[1]):
            for i in range(M.shape[0]):
                ct, seed = _seeded_encrypt(S, np.full(S.n // 2, M[i, j]), scale, L, rng)
                req += ref.save_ciphertext(S.n, S.moduli, ct, scale, seed, 1)
    res = _replies(S, _serve(tmp_path, "serve_batch_matmul", req, [2, 3, 3, 2]))
    assert len(res) == 4
    Cm = A @ B
    for j in range(2):
        for i in range(2):
            c = res[i + 2 * j]
            assert abs(S.decrypt(c["data"], c["scale"])[0].real - Cm[i, j]) < 1e-4
    # fft over 4 ciphertexts (client_side_fft's shape, shorter: two levels fit this chain): slot s of ciphertext i holds x_i(s)
    m = 4
    x = rng.uniform(-1, 1, m) + 1j * rng.uniform(-1, 1, m)
    req = ref.save_parms(S.n, S.moduli, 0)
    for i in range(m):
        ct, seed = _seeded_encrypt(S, np.full(S.n // 2, x[i]), scale, L, rng)
        req += ref.save_ciphertext(S.n, S.moduli, ct, scale, seed, 0)
    res = _replies(S, _serve(tmp_path, "serve_fft", req, [m]))
    got = np.array([S.decrypt(c["data"], c["scale"])[0] for c in res])
    assert np.max(np.abs(got - np.fft.fft(x))) < 1e-3, got
