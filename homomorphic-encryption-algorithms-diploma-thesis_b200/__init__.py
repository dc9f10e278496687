"""B200-native CKKS evaluator for the encrypted linear-algebra hot path of
isteiakakis/Homomorphic-Encryption-Algorithms-Diploma-Thesis.

This package is a thin ctypes binding of libhegpu.so (C ABI: include/hegpu.h).  All
arithmetic runs in hand-written sm_100a kernels (csrc/); there is no CPU fallback: if
the library or a GPU is missing, loading / context creation raises.

The directory name contains hyphens, so import it through `hegpu_loader.load()`
(repo root) or importlib; the module registers itself as `hegpu_b200`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhegpu.so")

u64p = C.POINTER(C.c_uint64)


class HegpuError(RuntimeError):
    """Base class; .status is the hegpu_status code."""

    status = 0


class InvalidArgument(HegpuError, ValueError):  # SEAL: std::invalid_argument
    status = 1


class LogicError(HegpuError):  # SEAL: std::logic_error
    status = 2


class CudaError(HegpuError):
    status = 3


class OutOfMemory(HegpuError, MemoryError):
    status = 4


_ERR = {1: InvalidArgument, 2: LogicError, 3: CudaError, 4: OutOfMemory}

_lib = None


def lib() -> C.CDLL:
    """Load libhegpu.so.  Fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.hegpu_last_error.restype = C.c_char_p
        L.hegpu_version.restype = C.c_char_p
        L.hegpu_ctx_stream.restype = C.c_void_p
        L.hegpu_ctx_stream.argtypes = [C.c_void_p]
        L.hegpu_ctx_psi.restype = C.c_uint64
        L.hegpu_ctx_psi.argtypes = [C.c_void_p, C.c_uint32]
        L.hegpu_launch_count.restype = C.c_uint64
        L.hegpu_launch_count.argtypes = [C.c_void_p]
        L.hegpu_profile_kind_name.restype = C.c_char_p
        L.hegpu_profile_kind_name.argtypes = [C.c_int]
        vp, u32, i32, dbl = C.c_void_p, C.c_uint32, C.c_int, C.c_double
        sigs = {
            "hegpu_ctx_create": [C.POINTER(vp), u32, u64p, u32, i32],
            "hegpu_ctx_destroy": [vp],
            "hegpu_sync": [vp],
            "hegpu_load_relin_key": [vp, vp],
            "hegpu_load_galois_key": [vp, u32, vp],
            "hegpu_has_galois_key": [vp, u32],
            "hegpu_galois_elt_from_step": [vp, i32, C.POINTER(u32)],
            "hegpu_ct_create": [vp, C.POINTER(vp), u32, u32, u32],
            "hegpu_ct_destroy": [vp],
            "hegpu_ct_upload": [vp, vp, u32, u32, dbl],
            "hegpu_ct_download": [vp, vp],
            "hegpu_ct_upload_one": [vp, u32, vp],
            "hegpu_ct_download_one": [vp, u32, vp],
            "hegpu_ct_upload_async": [vp, vp, u32, u32, dbl],
            "hegpu_ct_download_async": [vp, vp],
            "hegpu_ct_copy_wait": [vp],
            "hegpu_ct_info": [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(dbl)],
            "hegpu_ct_set_scale": [vp, dbl],
            "hegpu_ct_set_meta": [vp, u32, u32, dbl],
            "hegpu_ct_copy": [vp, vp, vp],
            "hegpu_ct_copy_one": [vp, vp, u32, vp, u32],
            "hegpu_ct_device_view": [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)],
            "hegpu_pt_create": [vp, C.POINTER(vp), u32, u32],
            "hegpu_pt_destroy": [vp],
            "hegpu_pt_upload": [vp, vp, u32, dbl],
            "hegpu_pt_upload_ext": [vp, vp, u32, dbl],
            "hegpu_pt_upload_one": [vp, u32, vp],
            "hegpu_pt_download_one": [vp, u32, vp],
            "hegpu_negate": [vp, vp, vp],
            "hegpu_add": [vp, vp, vp, vp],
            "hegpu_sub": [vp, vp, vp, vp],
            "hegpu_add_plain": [vp, vp, vp, vp, i32],
            "hegpu_sub_plain": [vp, vp, vp, vp, i32],
            "hegpu_multiply_plain": [vp, vp, vp, vp, i32],
            "hegpu_multiply": [vp, vp, vp, vp],
            "hegpu_square": [vp, vp, vp],
            "hegpu_relinearize": [vp, vp, vp],
            "hegpu_rescale_to_next": [vp, vp, vp],
            "hegpu_mod_switch_to_next": [vp, vp, vp],
            "hegpu_rotate_vector": [vp, vp, vp, i32],
            "hegpu_apply_galois": [vp, vp, vp, u32],
            "hegpu_ntt_forward_device": [vp, vp, u32, u32, u32],
            "hegpu_ntt_inverse_device": [vp, vp, u32, u32, u32],
            "hegpu_ntt_forward_host": [vp, vp, u32, u32, u32],
            "hegpu_ntt_inverse_host": [vp, vp, u32, u32, u32],
            "hegpu_matvec_bsgs": [vp, vp, vp, vp, u32, u32, i32],
            "hegpu_matvec_bsgs_range": [vp, vp, vp, vp, u32, u32, u32, i32],
            "hegpu_bmatmul": [vp, vp, vp, vp, u32, u32, i32],
            "hegpu_matmul_elemwise": [vp, vp, vp, vp, u32, u32, u32, i32, i32],
            "hegpu_bfft_stage": [vp, vp, vp, i32, i32],
            "hegpu_fft_butterflies": [vp, vp, vp, vp, vp, vp],
            "hegpu_reduce_fixup": [vp, vp, u32],
            "hegpu_rescale_sum_to_next": [vp, vp, vp, u32],
            "hegpu_ct_transparent": [vp, vp, C.POINTER(u32)],
            "hegpu_pipe_peak": [vp, i32, C.POINTER(dbl)],
            "hegpu_profile_enable": [vp, i32],
            "hegpu_profile_reset": [vp],
            "hegpu_profile_read": [vp, i32, C.POINTER(dbl), u64p, u64p, u64p],
        }
        for name, args in sigs.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        _lib = L
    return _lib


EXPORTS = None  # filled lazily by exported_symbols()


def _ck(status: int):
    if status:
        msg = lib().hegpu_last_error().decode()
        raise _ERR.get(status, HegpuError)(msg)


def _hp(a):
    """host pointer of a numpy uint64 array or a raw integer address"""
    if isinstance(a, np.ndarray):
        assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"], "need C-contiguous uint64"
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(int(a))


class Context:
    """seal::SEALContext + seal::Evaluator state on one GPU (hegpu_ctx)."""

    def __init__(self, n: int, moduli, device: int = 0):
        self.n = int(n)
        self.moduli = [int(q) for q in moduli]
        self.K = len(self.moduli)
        arr = (C.c_uint64 * self.K)(*self.moduli)
        h = C.c_void_p()
        _ck(lib().hegpu_ctx_create(C.byref(h), self.n, arr, self.K, device))
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib().hegpu_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- misc
    def sync(self):
        _ck(lib().hegpu_sync(self._h))

    @property
    def stream(self) -> int:
        return int(lib().hegpu_ctx_stream(self._h) or 0)

    def psi(self, i: int) -> int:
        return int(lib().hegpu_ctx_psi(self._h, i))

    @property
    def launches(self) -> int:
        return int(lib().hegpu_launch_count(self._h))

    def galois_elt_from_step(self, step: int) -> int:
        e = C.c_uint32()
        _ck(lib().hegpu_galois_elt_from_step(self._h, step, C.byref(e)))
        return e.value

    # ---- keys
    def load_relin_key(self, key):
        _ck(lib().hegpu_load_relin_key(self._h, _hp(key)))

    def load_galois_key(self, elt: int, key):
        _ck(lib().hegpu_load_galois_key(self._h, elt, _hp(key)))

    def load_galois_keys(self, keys: dict):
        for elt, k in keys.items():
            self.load_galois_key(int(elt), k)

    def has_galois_key(self, elt: int) -> bool:
        return bool(lib().hegpu_has_galois_key(self._h, elt))

    # ---- factories
    def ct(self, batch: int, size_cap: int = 3, L_cap: int | None = None) -> "CtBatch":
        return CtBatch(self, batch, size_cap, self.K - 1 if L_cap is None else L_cap)

    def pt(self, count: int, L_cap: int | None = None) -> "PtSet":
        return PtSet(self, count, self.K - 1 if L_cap is None else L_cap)

    def upload_ct(self, host: np.ndarray, scale: float, size_cap: int = 3, L_cap: int | None = None) -> "CtBatch":
        """host [B][size][L][N] (or [size][L][N] for a batch of one)."""
        if host.ndim == 3:
            host = host[None]
        host = np.ascontiguousarray(host)
        t = self.ct(host.shape[0], max(size_cap, host.shape[1]), L_cap)
        t.upload(host, scale)
        return t

    def upload_pt(self, host: np.ndarray, scale: float, L_cap: int | None = None) -> "PtSet":
        if host.ndim == 2:
            host = host[None]
        host = np.ascontiguousarray(host)
        t = self.pt(host.shape[0], L_cap)
        t.upload(host, scale)
        return t

    def upload_pt_ext(self, host: np.ndarray, scale: float) -> "PtSet":
        """host [count][L+1][N]: L data limbs + one limb mod the special prime (plaintexts of the
        double-hoisted matvec, hegpu_pt_upload_ext)."""
        if host.ndim == 2:
            host = host[None]
        host = np.ascontiguousarray(host)
        t = self.pt(host.shape[0], host.shape[1])
        t.upload_ext(host, scale)
        return t

    # ---- NTT (measurement / tooling)
    def ntt_forward_host(self, a: np.ndarray, first_mod: int, n_mods: int = 1):
        _ck(lib().hegpu_ntt_forward_host(self._h, _hp(a), a.size // self.n, first_mod, n_mods))

    def ntt_inverse_host(self, a: np.ndarray, first_mod: int, n_mods: int = 1):
        _ck(lib().hegpu_ntt_inverse_host(self._h, _hp(a), a.size // self.n, first_mod, n_mods))

    def ntt_forward_device(self, dptr: int, count: int, first_mod: int, n_mods: int):
        _ck(lib().hegpu_ntt_forward_device(self._h, C.c_void_p(dptr), count, first_mod, n_mods))

    def ntt_inverse_device(self, dptr: int, count: int, first_mod: int, n_mods: int):
        _ck(lib().hegpu_ntt_inverse_device(self._h, C.c_void_p(dptr), count, first_mod, n_mods))

    def pipe_peak(self, kind: int) -> float:
        """Measured issue rate (thread-level ops/s) of IMAD.WIDE.U32 (0), DFMA (1) or IMAD (2) on this device."""
        v = C.c_double()
        _ck(lib().hegpu_pipe_peak(self._h, kind, C.byref(v)))
        return v.value

    # ---- evaluator (out may alias a)
    def negate(self, out, a):
        _ck(lib().hegpu_negate(self._h, out._h, a._h))

    def add(self, out, a, b):
        _ck(lib().hegpu_add(self._h, out._h, a._h, b._h))

    def sub(self, out, a, b):
        _ck(lib().hegpu_sub(self._h, out._h, a._h, b._h))

    def add_plain(self, out, a, pt, index: int = 0):
        _ck(lib().hegpu_add_plain(self._h, out._h, a._h, pt._h, index))

    def sub_plain(self, out, a, pt, index: int = 0):
        _ck(lib().hegpu_sub_plain(self._h, out._h, a._h, pt._h, index))

    def multiply_plain(self, out, a, pt, index: int = 0):
        _ck(lib().hegpu_multiply_plain(self._h, out._h, a._h, pt._h, index))

    def multiply(self, out, a, b):
        _ck(lib().hegpu_multiply(self._h, out._h, a._h, b._h))

    def square(self, out, a):
        _ck(lib().hegpu_square(self._h, out._h, a._h))

    def relinearize(self, out, a):
        _ck(lib().hegpu_relinearize(self._h, out._h, a._h))

    def rescale_to_next(self, out, a):
        _ck(lib().hegpu_rescale_to_next(self._h, out._h, a._h))

    def mod_switch_to_next(self, out, a):
        _ck(lib().hegpu_mod_switch_to_next(self._h, out._h, a._h))

    def rotate_vector(self, out, a, steps: int):
        _ck(lib().hegpu_rotate_vector(self._h, out._h, a._h, steps))

    def apply_galois(self, out, a, elt: int):
        _ck(lib().hegpu_apply_galois(self._h, out._h, a._h, elt))

    # ---- composites
    def matvec_bsgs(self, out, a, diags, n1: int, n2: int, rescale: bool = True, hoist: bool = False, lazy: bool | None = None,
                    g_first: int = 0, dh: bool = False, imma: bool = False):
        """hoist: HEGPU_MATVEC_HOIST (hoisted baby steps); lazy: HEGPU_MATVEC_LAZY (one mod-down for all giant
        steps; defaults to `hoist`); dh: HEGPU_MATVEC_DH (double-hoisted, diags uploaded with upload_pt_ext);
        g_first: first global giant step of a diagonal-sharded call."""
        lazy = hoist if lazy is None else lazy
        flags = (1 if rescale else 0) | (2 if hoist else 0) | (4 if lazy else 0) | (8 if dh else 0) | (16 if imma else 0)
        _ck(lib().hegpu_matvec_bsgs_range(self._h, out._h, a._h, diags._h, n1, n2, g_first, flags))

    def bmatmul(self, out, this_cts, other_cts, n: int, p: int, case_b: bool):
        _ck(lib().hegpu_bmatmul(self._h, out._h, this_cts._h, other_cts._h, n, p, int(case_b)))

    def matmul_elemwise(self, out, a, b, rows: int, inner: int, cols: int, a_t: bool = False, b_t: bool = False):
        _ck(lib().hegpu_matmul_elemwise(self._h, out._h, a._h, b._h, rows, inner, cols, int(a_t), int(b_t)))

    def bfft_stage(self, y, stage_pts, steps: int, with_d2: bool):
        _ck(lib().hegpu_bfft_stage(self._h, y._h, stage_pts._h, steps, int(with_d2)))

    def fft_butterflies(self, out, even, odd, w_pts, one_pt):
        _ck(lib().hegpu_fft_butterflies(self._h, out._h, even._h, odd._h, w_pts._h, one_pt._h))

    # ---- per-kernel timing (bench roofline)
    def profile(self, on: bool):
        _ck(lib().hegpu_profile_enable(self._h, int(on)))

    def profile_reset(self):
        _ck(lib().hegpu_profile_reset(self._h))

    def profile_read(self) -> dict:
        out = {}
        k = 0
        while True:
            name = lib().hegpu_profile_kind_name(k)
            if name is None:
                break
            ms, la, un, by = C.c_double(), C.c_uint64(), C.c_uint64(), C.c_uint64()
            _ck(lib().hegpu_profile_read(self._h, k, C.byref(ms), C.byref(la), C.byref(un), C.byref(by)))
            out[name.decode()] = {"ms": ms.value, "launches": la.value, "units": un.value, "algo_bytes": by.value}
            k += 1
        return out

    def reduce_fixup(self, ct, terms: int):
        _ck(lib().hegpu_reduce_fixup(self._h, ct._h, terms))

    def rescale_sum_to_next(self, out, a, terms: int):
        """rescale_to_next of the uint64 sum of `terms` partial ciphertexts (the fix-up rides in the rescale's loads)."""
        _ck(lib().hegpu_rescale_sum_to_next(self._h, out._h, a._h, terms))

    def transparent_count(self, ct) -> int:
        """Number of transparent ciphertexts of the batch (Ciphertext::is_transparent); blocking."""
        n = C.c_uint32(0)
        _ck(lib().hegpu_ct_transparent(self._h, ct._h, C.byref(n)))
        return int(n.value)


class CtBatch:
    """A batch of seal::Ciphertext resident in HBM (hegpu_ct)."""

    def __init__(self, ctx: Context, batch: int, size_cap: int, L_cap: int):
        self.ctx = ctx
        h = C.c_void_p()
        _ck(lib().hegpu_ct_create(ctx._h, C.byref(h), batch, size_cap, L_cap))
        self._h = h
        self.batch, self.size_cap, self.L_cap = batch, size_cap, L_cap

    def __del__(self):
        try:
            if self._h and self.ctx._h:
                lib().hegpu_ct_destroy(self._h)
            self._h = None
        except Exception:
            pass

    def info(self):
        b, s, L, sc = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_double()
        _ck(lib().hegpu_ct_info(self._h, C.byref(b), C.byref(s), C.byref(L), C.byref(sc)))
        return b.value, s.value, L.value, sc.value

    @property
    def size(self):
        return self.info()[1]

    @property
    def L(self):
        return self.info()[2]

    @property
    def scale(self):
        return self.info()[3]

    @scale.setter
    def scale(self, v: float):
        _ck(lib().hegpu_ct_set_scale(self._h, float(v)))

    def upload(self, host, scale: float, size: int | None = None, L: int | None = None):
        """host: numpy [B][size][L][N] or a raw (pinned) host address with explicit size/L."""
        if isinstance(host, np.ndarray):
            assert host.shape[0] == self.batch and host.shape[3] == self.ctx.n
            size, L = host.shape[1], host.shape[2]
        _ck(lib().hegpu_ct_upload(self._h, _hp(host), size, L, float(scale)))

    def download(self, out=None) -> np.ndarray:
        b, s, L, _ = self.info()
        if out is None:
            out = np.empty((b, s, L, self.ctx.n), dtype=np.uint64)
        _ck(lib().hegpu_ct_download(self._h, _hp(out)))
        return out

    def upload_async(self, host_ptr: int, scale: float, size: int, L: int):
        """asynchronous H2D from pinned host memory (raw address); see hegpu_ct_upload_async"""
        _ck(lib().hegpu_ct_upload_async(self._h, C.c_void_p(int(host_ptr)), size, L, float(scale)))

    def download_async(self, host_ptr: int):
        _ck(lib().hegpu_ct_download_async(self._h, C.c_void_p(int(host_ptr))))

    def copy_wait(self):
        _ck(lib().hegpu_ct_copy_wait(self._h))

    def upload_one(self, index: int, host: np.ndarray):
        _ck(lib().hegpu_ct_upload_one(self._h, index, _hp(host)))

    def download_one(self, index: int) -> np.ndarray:
        _, s, L, _ = self.info()
        out = np.empty((s, L, self.ctx.n), dtype=np.uint64)
        _ck(lib().hegpu_ct_download_one(self._h, index, _hp(out)))
        return out

    def copy_from(self, src: "CtBatch"):
        _ck(lib().hegpu_ct_copy(self.ctx._h, self._h, src._h))

    def set_meta(self, size: int, L: int, scale: float):
        """Declare the contents after the device buffer was filled externally (NCCL receive)."""
        _ck(lib().hegpu_ct_set_meta(self._h, size, L, float(scale)))

    def device_view(self):
        p, sb, sp, sl = C.c_void_p(), C.c_size_t(), C.c_size_t(), C.c_size_t()
        _ck(lib().hegpu_ct_device_view(self._h, C.byref(p), C.byref(sb), C.byref(sp), C.byref(sl)))
        return int(p.value), sb.value, sp.value, sl.value


class PtSet:
    """A set of seal::Plaintext (NTT form) resident in HBM (hegpu_pt)."""

    def __init__(self, ctx: Context, count: int, L_cap: int):
        self.ctx = ctx
        h = C.c_void_p()
        _ck(lib().hegpu_pt_create(ctx._h, C.byref(h), count, L_cap))
        self._h = h
        self.count, self.L_cap = count, L_cap

    def __del__(self):
        try:
            if self._h and self.ctx._h:
                lib().hegpu_pt_destroy(self._h)
            self._h = None
        except Exception:
            pass

    def upload(self, host: np.ndarray, scale: float):
        assert host.shape[0] == self.count and host.shape[2] == self.ctx.n
        _ck(lib().hegpu_pt_upload(self._h, _hp(host), host.shape[1], float(scale)))

    def upload_ext(self, host: np.ndarray, scale: float):
        """host [count][L+1][N], limb L = residues mod the special prime"""
        assert host.shape[0] == self.count and host.shape[2] == self.ctx.n
        _ck(lib().hegpu_pt_upload_ext(self._h, _hp(host), host.shape[1] - 1, float(scale)))

    def upload_one(self, index: int, host: np.ndarray):
        _ck(lib().hegpu_pt_upload_one(self._h, index, _hp(host)))
