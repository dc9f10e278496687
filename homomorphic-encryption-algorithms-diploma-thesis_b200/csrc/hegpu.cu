// hegpu.cu -- host side of libhegpu.so: context/tables, device handles, the C ABI of
// include/hegpu.h and the composites that stay on the device.  No CPU arithmetic on
// ciphertext data happens here; the host only builds constant tables and launches kernels.
#include "ntt_launch.cuh"
#include "mac_kernels.cuh"
#include "imma_kernels.cuh"

// ------------------------------------------------------------------------- errors
static thread_local std::string g_err;
int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

extern "C" const char *hegpu_last_error(void) { return g_err.c_str(); }
extern "C" const char *hegpu_version(void) { return "hegpu 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------- host number theory
typedef unsigned __int128 u128;
static u64 h_mulmod(u64 a, u64 b, u64 q) { return (u64)(((u128)a * b) % q); }
static u64 h_powmod(u64 a, u64 e, u64 q)
{
    u64 r = 1;
    a %= q;
    while (e) {
        if (e & 1) r = h_mulmod(r, a, q);
        a = h_mulmod(a, a, q);
        e >>= 1;
    }
    return r;
}
static u64 h_invmod(u64 a, u64 q) { return h_powmod(a, q - 2, q); }
static u64 h_shoup(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }
static bool h_is_prime(u64 n)
{
    if (n < 2) return false;
    static const u64 sp[] = { 2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37 };
    for (u64 p : sp) {
        if (n == p) return true;
        if (n % p == 0) return false;
    }
    u64 d = n - 1;
    int r = 0;
    while (!(d & 1)) { d >>= 1; ++r; }
    for (u64 a : sp) {
        u64 x = h_powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int j = 1; j < r; ++j) {
            x = h_mulmod(x, x, n);
            if (x == n - 1) { comp = false; break; }
        }
        if (comp) return false;
    }
    return true;
}
static u32 h_brev(u32 x, u32 bits)
{
    u32 r = 0;
    for (u32 i = 0; i < bits; ++i) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}
// the numerically smallest primitive 2N-th root of unity (SEAL try_minimal_primitive_root)
static u64 h_min_root(u64 q, u32 n)
{
    const u64 e = (q - 1) / (2ull * n);
    u64 root = 0;
    for (u64 g = 2;; ++g) {
        u64 r = h_powmod(g, e, q);
        if (h_powmod(r, n, q) == q - 1) { root = r; break; }
    }
    const u64 sq = h_mulmod(root, root, q);
    u64 cur = root, best = root;
    for (u32 i = 0; i < n; ++i) {
        best = std::min(best, cur);
        cur = h_mulmod(cur, sq, q);
    }
    return best;
}

int set_device(hegpu_ctx *c)
{
    CU(cudaSetDevice(c->device));
    return HEGPU_OK;
}

int arena_reserve(hegpu_ctx *c, size_t bytes)
{
    if (bytes <= c->arena.cap) return HEGPU_OK;
    CU(cudaStreamSynchronize(c->stream));
    if (c->arena.base) CU(cudaFree(c->arena.base));
    c->arena.base = nullptr;
    c->arena.cap = 0;
    size_t want = bytes + (bytes >> 3);
    CU(cudaMalloc(&c->arena.base, want));
    c->arena.cap = want;
    return HEGPU_OK;
}

int park_reserve(hegpu_ctx *c, size_t words)
{
    if (words <= c->park_words) return HEGPU_OK;
    CU(cudaStreamSynchronize(c->stream));
    if (c->park) CU(cudaFree(c->park));
    c->park = nullptr;
    c->park_words = 0;
    words += words >> 2;
    CU(cudaMalloc(&c->park, words * sizeof(u64)));
    c->park_words = words;
    return HEGPU_OK;
}

void select_slot(hegpu_ctx *c, int slot)
{
    c->park_slots[c->cur_slot] = c->park;
    c->park_words_slots[c->cur_slot] = c->park_words;
    c->cur_slot = slot;
    c->park = c->park_slots[slot];
    c->park_words = c->park_words_slots[slot];
    c->stream = slot ? c->aux_stream[slot] : c->main_stream;
}
struct SlotGuard {  // composites return to slot 0 on every exit path, and the main stream waits for every
    hegpu_ctx *c;   // auxiliary stream that was forked off it (also on an early error return)
    int forked = 0;
    ~SlotGuard()
    {
        select_slot(c, 0);
        for (int sl = 1; sl < forked; ++sl)
            if (cudaEventRecord(c->ev_join[sl], c->aux_stream[sl]) == cudaSuccess) cudaStreamWaitEvent(c->main_stream, c->ev_join[sl], 0);
    }
};

int stage_reserve(hegpu_ctx *c, size_t words)
{
    if (words <= c->stage_words) return HEGPU_OK;
    CU(cudaStreamSynchronize(c->stream));
    if (c->stage) CU(cudaFree(c->stage));
    c->stage = nullptr;
    c->stage_words = 0;
    CU(cudaMalloc(&c->stage, words * sizeof(u64)));
    c->stage_words = words;
    return HEGPU_OK;
}

int configure_smem(hegpu_ctx *c, const void *kernel, size_t bytes)
{
    auto it = c->smem_configured.find(kernel);
    if (it != c->smem_configured.end() && it->second >= bytes) return HEGPU_OK;
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    c->smem_configured[kernel] = bytes;
    return HEGPU_OK;
}

int ew_grid(hegpu_ctx *c, size_t total)
{
    size_t blocks = (total + 255) / 256;
    size_t cap = (size_t)c->sms * 16;
    return (int)std::max<size_t>(1, std::min(blocks, cap));
}

// ------------------------------------------------------------------------- context
extern "C" int hegpu_ctx_create(hegpu_ctx **out, uint32_t n, const uint64_t *moduli, uint32_t K, int device)
{
    if (!out || !moduli) INVALID("null argument");
    if (n != 4096 && n != 8192 && n != 16384 && n != 32768) INVALID("poly_modulus_degree must be 4096, 8192, 16384 or 32768");
    if (K < 1 || K > 62) INVALID("coeff_modulus size out of range");
    {   // the key inner product sums K-1 products below q^2 lazily and reduces once: needs (K-1) * q < 2^64
        u64 mx = 0;
        for (u32 i = 0; i < K; ++i) mx = std::max<u64>(mx, moduli[i]);
        if (K > 1 && mx > (~0ull) / (K - 1)) INVALID("coeff_modulus too large: (size - 1) * max prime must stay below 2^64");
    }
    int ndev = 0;
    cudaError_t de = cudaGetDeviceCount(&ndev);
    if (de != cudaSuccess || ndev == 0)
        return fail(HEGPU_ERR_CUDA, "no CUDA device: libhegpu has no CPU fallback");
    if (device < 0 || device >= ndev) INVALID("invalid device ordinal");
    for (u32 i = 0; i < K; ++i) {
        if (moduli[i] >> 60) INVALID("coeff_modulus primes must be at most 60 bits");
        if (!h_is_prime(moduli[i]) || (moduli[i] - 1) % (2ull * n)) INVALID("coeff_modulus primes must be prime and = 1 mod 2N");
        for (u32 j = 0; j < i; ++j)
            if (moduli[i] == moduli[j]) INVALID("coeff_modulus primes must be distinct");
    }
    hegpu_ctx *c = new hegpu_ctx();
    c->device = device;
    c->n = n;
    c->K = K;
    while ((1u << c->logn) < n) c->logn++;
    c->q.assign(moduli, moduli + K);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->sms = prop.multiProcessorCount;
    if (const char *e = getenv("HEGPU_LOGE")) c->loge = atoi(e) == 3 ? 3 : 4;
    if (const char *e = getenv("HEGPU_PARK")) c->use_park = atoi(e) != 0;
    if (const char *e = getenv("HEGPU_DH_FUSED")) c->dh_fused = atoi(e) != 0;
    if (const char *e = getenv("HEGPU_DH_F64")) c->dh_f64 = atoi(e) != 0;
    if (const char *e = getenv("HEGPU_DH_SWZ")) c->dh_swz = atoi(e) != 0;
    if (const char *e = getenv("HEGPU_DH_STCS")) c->dh_stcs = atoi(e) != 0;
    if (const char *e = getenv("HEGPU_DH_BK")) c->dh_bk = atoi(e) != 0;
    if (const char *e = getenv("HEGPU_LIMB_MAJOR")) c->limb_major = atoi(e) != 0;
    if (const char *e = getenv("HEGPU_DH_IMMA")) c->dh_imma = atoi(e) != 0;
    if (const char *e = getenv("HEGPU_IMMA_TX")) c->imma_tx = atoi(e) == 16 ? 16 : 8;
    if (const char *e = getenv("HEGPU_FUSE_FINAL")) c->fuse_final = atoi(e) != 0;
    if (const char *e = getenv("HEGPU_PARK32K")) c->park32k = atoi(e) < 0 ? 0 : atoi(e) > 2 ? 2 : atoi(e);
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->main_stream = c->stream;
    CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    for (int sl = 1; sl < hegpu_ctx::MAX_SLOTS; ++sl) {
        CU(cudaStreamCreateWithFlags(&c->aux_stream[sl], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->ev_join[sl], cudaEventDisableTiming));
    }
    if (const char *e = getenv("HEGPU_STREAMS")) c->n_slots = std::min(std::max(atoi(e), 1), (int)hegpu_ctx::MAX_SLOTS);
    CU(cudaStreamCreateWithFlags(&c->copy_h2d, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->copy_d2h, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev_fence, cudaEventDisableTiming));

    // bit counts of q_0..q_{L-1} products (SEAL total_coeff_modulus_bit_count)
    {
        std::vector<u64> big(1, 1);
        for (u32 i = 0; i < K; ++i) {
            u64 carry = 0;
            for (auto &w : big) {
                u128 t = (u128)w * c->q[i] + carry;
                w = (u64)t;
                carry = (u64)(t >> 64);
            }
            if (carry) big.push_back(carry);
            int bits = (int)(big.size() - 1) * 64 + (64 - __builtin_clzll(big.back()));
            c->level_bits.push_back(bits);
        }
    }

    std::vector<ulonglong2> fwd((size_t)K * n), inv((size_t)K * n), inv_last(K);
    std::vector<ModConst> mods(K);
    std::vector<MdConst> md((size_t)K * K);
    std::vector<double> fwd_d((size_t)K * n), inv_d((size_t)K * n);
    std::vector<ModF64> modsd(K);
    for (u32 i = 0; i < K; ++i) {
        const u64 q = c->q[i];
        const u64 psi = h_min_root(q, n), ipsi = h_invmod(psi, q);
        c->psi.push_back(psi);
        u64 p = 1, ip = 1;
        for (u32 k = 0; k < n; ++k) {
            const u32 r = h_brev(k, c->logn);
            fwd[(size_t)i * n + r] = make_ulonglong2(p, h_shoup(p, q));
            inv[(size_t)i * n + r] = make_ulonglong2(ip, h_shoup(ip, q));
            fwd_d[(size_t)i * n + r] = (double)p;  // exact below 2^53; only used when q < 2^43
            inv_d[(size_t)i * n + r] = (double)ip;
            p = h_mulmod(p, psi, q);
            ip = h_mulmod(ip, ipsi, q);
        }
        ModConst &m = mods[i];
        m.q = q;
        u128 mu = (~(u128)0) / q;  // floor((2^128-1)/q) == floor(2^128/q) for odd q > 1
        m.mu_hi = (u64)(mu >> 64);
        m.mu_lo = (u64)mu;
        m.ninv = h_invmod(n % q, q);
        m.ninv_sh = h_shoup(m.ninv, q);
        // bit 0: forward needs corrections; bit 1: inverse needs corrections; bit 2: FP64 butterflies
        m.big = (q >> 58 ? 1u : 0u) | (q >> 46 ? 2u : 0u) | ((q >> 43) == 0 && !getenv("HEGPU_NO_FP64") ? 4u : 0u);
        m.pad = 0;
        m.q3 = q * 3;
        {   // -q^-1 mod 2^64 by Newton iteration; 2^64 mod q
            u64 inv = q;  // q*q = 1 mod 8
            for (int it = 0; it < 6; ++it) inv *= 2 - q * inv;
            m.qinv_neg = (u64)0 - inv;
            m.rmod = (u64)((((u128)1) << 64) % q);
            m.rmod_sh = h_shoup(m.rmod, q);
            m.pmont = (i + 1 < K) ? h_mulmod(c->q[K - 1] % q, m.rmod, q) : 0;
        }
        const u64 wl = h_mulmod(inv[(size_t)i * n + 1].x, m.ninv, q);
        inv_last[i] = make_ulonglong2(wl, h_shoup(wl, q));
        modsd[i] = ModF64{ (double)q, 1.0 / (double)q, (double)m.ninv, (double)wl };
        for (u32 d = 0; d < K; ++d) {
            MdConst &e = md[(size_t)d * K + i];
            e.pad = 0;
            if (d == i) { e.inv = e.inv_sh = e.halfmod = 0; continue; }
            e.inv = h_invmod(c->q[d] % q, q);
            e.inv_sh = h_shoup(e.inv, q);
            e.halfmod = (c->q[d] >> 1) % q;
        }
    }
    CU(cudaMalloc(&c->d_fwd, fwd.size() * sizeof(ulonglong2)));
    CU(cudaMalloc(&c->d_inv, inv.size() * sizeof(ulonglong2)));
    CU(cudaMalloc(&c->d_inv_last, inv_last.size() * sizeof(ulonglong2)));
    CU(cudaMalloc(&c->d_mods, mods.size() * sizeof(ModConst)));
    CU(cudaMalloc(&c->d_md, md.size() * sizeof(MdConst)));
    CU(cudaMalloc(&c->d_fwd_d, fwd_d.size() * sizeof(double)));
    CU(cudaMalloc(&c->d_inv_d, inv_d.size() * sizeof(double)));
    CU(cudaMalloc(&c->d_modsd, modsd.size() * sizeof(ModF64)));
    CU(cudaMemcpy(c->d_fwd_d, fwd_d.data(), fwd_d.size() * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_inv_d, inv_d.data(), inv_d.size() * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_modsd, modsd.data(), modsd.size() * sizeof(ModF64), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_fwd, fwd.data(), fwd.size() * sizeof(ulonglong2), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_inv, inv.data(), inv.size() * sizeof(ulonglong2), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_inv_last, inv_last.data(), inv_last.size() * sizeof(ulonglong2), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_mods, mods.data(), mods.size() * sizeof(ModConst), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_md, md.data(), md.size() * sizeof(MdConst), cudaMemcpyHostToDevice));
    c->tabs.fwd = c->d_fwd;
    c->tabs.inv = c->d_inv;
    c->tabs.inv_last = c->d_inv_last;
    c->tabs.mods = c->d_mods;
    c->tabs.fwd_d = c->d_fwd_d;
    c->tabs.inv_d = c->d_inv_d;
    c->tabs.modsd = c->d_modsd;
    c->tabs.n = n;
    c->tabs.logn = c->logn;
    *out = c;
    return HEGPU_OK;
}

extern "C" int hegpu_ctx_destroy(hegpu_ctx *c)
{
    if (!c) return HEGPU_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->d_fwd);
    cudaFree(c->d_inv);
    cudaFree(c->d_inv_last);
    cudaFree(c->d_mods);
    cudaFree(c->d_md);
    cudaFree(c->d_fwd_d);
    cudaFree(c->d_inv_d);
    cudaFree(c->d_modsd);
    cudaFree(c->relin_key);
    for (auto &kv : c->galois_keys) cudaFree(kv.second);
    for (auto &kv : c->perms) cudaFree(kv.second);
    for (auto &kv : c->tmp_cache) {
        hegpu_ct *t = kv.second;
        kv.second = nullptr;
        hegpu_ct_destroy(t);
    }
    cudaFree(c->arena.base);
    cudaFree(c->stage);
    select_slot(c, 0);
    cudaFree(c->park);
    cudaEventDestroy(c->ev_fork);
    for (int sl = 1; sl < hegpu_ctx::MAX_SLOTS; ++sl) {
        cudaStreamSynchronize(c->aux_stream[sl]);
        cudaFree(c->park_slots[sl]);
        cudaStreamDestroy(c->aux_stream[sl]);
        cudaEventDestroy(c->ev_join[sl]);
    }
    cudaStreamSynchronize(c->copy_h2d);
    cudaStreamSynchronize(c->copy_d2h);
    cudaStreamDestroy(c->copy_h2d);
    cudaStreamDestroy(c->copy_d2h);
    cudaEventDestroy(c->ev_fence);
    cudaStreamDestroy(c->stream);
    delete c;
    return HEGPU_OK;
}

extern "C" int hegpu_sync(hegpu_ctx *c)
{
    if (!c) INVALID("null context");
    TRY(set_device(c));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->copy_h2d));
    CU(cudaStreamSynchronize(c->copy_d2h));
    return HEGPU_OK;
}
extern "C" void *hegpu_ctx_stream(hegpu_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" uint64_t hegpu_ctx_psi(hegpu_ctx *c, uint32_t i) { return (c && i < c->K) ? c->psi[i] : 0; }
extern "C" uint64_t hegpu_launch_count(hegpu_ctx *c) { return c ? c->launches : 0; }

// ------------------------------------------------------------------------- profiling
static int prof_collect(hegpu_ctx *c)
{
    if (c->prof.empty()) return HEGPU_OK;
    CU(cudaStreamSynchronize(c->stream));
    for (auto &r : c->prof) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, r.a, r.b));
        c->prof_ms[r.kind] += ms;
        c->prof_launches[r.kind]++;
        c->prof_units[r.kind] += r.units;
        c->prof_bytes[r.kind] += r.bytes;
        c->ev_pool.push_back(r.a);
        c->ev_pool.push_back(r.b);
    }
    c->prof.clear();
    return HEGPU_OK;
}
extern "C" int hegpu_profile_enable(hegpu_ctx *c, int on)
{
    if (!c) INVALID("null context");
    TRY(set_device(c));
    TRY(prof_collect(c));
    c->profiling = on != 0;
    return HEGPU_OK;
}
extern "C" int hegpu_profile_reset(hegpu_ctx *c)
{
    if (!c) INVALID("null context");
    TRY(set_device(c));
    TRY(prof_collect(c));
    for (int k = 0; k < PK_COUNT; ++k) {
        c->prof_ms[k] = 0;
        c->prof_launches[k] = c->prof_units[k] = c->prof_bytes[k] = 0;
    }
    return HEGPU_OK;
}
extern "C" const char *hegpu_profile_kind_name(int kind) { return (kind >= 0 && kind < PK_COUNT) ? kProfNames[kind] : nullptr; }
extern "C" int hegpu_profile_read(hegpu_ctx *c, int kind, double *ms, uint64_t *launches, uint64_t *units, uint64_t *bytes)
{
    if (!c) INVALID("null context");
    if (kind < 0 || kind >= PK_COUNT) INVALID("profile kind out of range");
    TRY(set_device(c));
    TRY(prof_collect(c));
    if (ms) *ms = c->prof_ms[kind];
    if (launches) *launches = c->prof_launches[kind];
    if (units) *units = c->prof_units[kind];
    if (bytes) *bytes = c->prof_bytes[kind];
    return HEGPU_OK;
}

// ------------------------------------------------------------------------- pipe peaks (measurement)
// Issue-rate microbenchmarks behind the integer / FP64 rooflines that bench.py reports: independent dependent-chains
// per thread of ONE instruction kind or ONE arithmetic sequence of the product kernels, enough CTAs to fill every SM.
//   kind 0: IMAD.WIDE.U32 (32 x 32 -> 64 multiply-add, the building block of every 64-bit modular product)
//   kind 1: DFMA
//   kind 2: IMAD (32-bit low multiply-add)
//   kind 3: mac128 (modarith.cuh): one 64 x 64 -> 128-bit multiply-accumulate, the unit of dh_inner / ks_inner
//   kind 4: the 60-bit forward NTT butterfly (ArI64<true>::fwd_bfly: Shoup product + add/sub) with the per-pass
//           range correction every third stage, as in the radix-8 passes
//   kind 5: the FP64 forward NTT butterfly (ArF64::fwd_bfly, primes below 2^43)
// Kinds 3-5 run the product's own device functions on register operands only: what the arithmetic alone allows
// when no load, exchange, address computation or barrier is in the way.
template <int KIND>
__global__ void __launch_bounds__(256) pipe_peak_kernel(u64 *__restrict__ out, u32 iters, u32 seed, u32 sink, const ModConst *mods,
                                                        const ModF64 *modsd)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    u32 x = t * 2654435761u + seed, y = (t ^ seed) | 1u;
    u64 acc = 0;
    if constexpr (KIND <= 2) {
        u64 a[8];
        double d[8];
        u32 r[8];
        const double dx = 1.0 + (double)(x & 1023u) * 1e-9, dy = 1e-9 * (double)(y & 1023u);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            a[i] = (u64)t + i;
            d[i] = (double)i;
            r[i] = t + i;
        }
        for (u32 it = 0; it < iters; ++it) {
#pragma unroll
            for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    // multiplicand = low word of the neighbouring chain: changes every iteration, so nothing can be hoisted
                    if (KIND == 0)
                        asm volatile("{\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %1;\n\tmad.wide.u32 %0, lo, %2, %0;\n\t}"
                                     : "+l"(a[i])
                                     : "l"(a[(i + 1) & 7]), "r"(y));
                    if (KIND == 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dx), "d"(dy));
                    if (KIND == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(x), "r"(y));
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += a[i] + (u64)__double_as_longlong(d[i]) + r[i];
    } else if constexpr (KIND == 3) {
        // 8 accumulators x 4 operand pairs per iteration = 32 products; operands below 2^62 in distinct registers
        u64 hi[8], lo[8], xs[8], ys[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            hi[i] = 0;
            lo[i] = t + i;
            xs[i] = (((u64)x << 29) ^ ((u64)y * (2 * i + 1))) & 0x3FFFFFFFFFFFFFFFull;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) ys[j] = (((u64)y << 30) ^ ((u64)x * (2 * j + 3))) & 0x3FFFFFFFFFFFFFFFull;
        for (u32 it = 0; it < iters; ++it) {
#pragma unroll
            for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
                for (int i = 0; i < 8; ++i) mac128(hi[i], lo[i], xs[(i + rep) & 7], ys[rep]);
            }
            // new operands every iteration (in the product kernels they come from loads): nothing loop-invariant to hoist
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                hi[i] &= 0xFFFFFFFull;  // keep the sums from wrapping
                xs[i] = (xs[i] + lo[(i + 3) & 7]) & 0x3FFFFFFFFFFFFFFFull;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) ys[j] = (ys[j] ^ lo[j]) & 0x3FFFFFFFFFFFFFFFull;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += hi[i] ^ lo[i];
    } else if constexpr (KIND == 4) {
        // 4 independent butterflies per stage (a radix-8 register set), 3 stages between range corrections, 32 per iteration
        const ModConst m = mods[0];
        const ArI64<true> ar(m, make_ulonglong2(0, 0));
        u64 v[8];
        ulonglong2 W[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = ((u64)x * (i + 1) + y) % m.q;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            W[j].x = ((u64)y * (j + 5) + x) % m.q;
            W[j].y = (u64)(((unsigned __int128)W[j].x << 64) / m.q);
        }
        for (u32 it = 0; it < iters; ++it) {
#pragma unroll
            for (int rep = 0; rep < 8; ++rep) {
                if (rep % 3 == 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = ar.fwd_fix(v[i]);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int s = rep % 3, lo_ = j & ((1 << s) - 1), k = ((j >> s) << (s + 1)) | lo_;
                    if (s < ArI64<true>::approx_stages(3))  // as in a radix-8 pass: the approximate quotient in two of three stages
                        ar.fwd_bfly<true>(v[k], v[k | (1 << s)], W[(j + rep) & 3]);
                    else
                        ar.fwd_bfly<false>(v[k], v[k | (1 << s)], W[(j + rep) & 3]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += v[i];
    } else {
        const ArF64 ar(modsd[1]);
        double v[8], W[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (double)((x * (i + 1) + y) & 0xFFFFFu);
#pragma unroll
        for (int j = 0; j < 4; ++j) W[j] = (double)((y * (j + 5) + x) & 0xFFFFFFFu);
        for (u32 it = 0; it < iters; ++it) {
#pragma unroll
            for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int s = rep % 3, lo_ = j & ((1 << s) - 1), k = ((j >> s) << (s + 1)) | lo_;
                    ar.fwd_bfly<false>(v[k], v[k | (1 << s)], W[(j + rep) & 3]);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = ar.reduce(v[i]);  // once per 8 stages, as the per-pass fix of the inverse transform
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += (u64)__double_as_longlong(v[i]);
    }
    if (sink) out[t] = acc;  // sink = 0 at run time; keeps the chains alive
}

extern "C" int hegpu_pipe_peak(hegpu_ctx *c, int kind, double *ops_per_second)
{
    if (!c || !ops_per_second) INVALID("null argument");
    if (kind < 0 || kind > 5) INVALID("pipe kind out of range");
    if (kind == 5 && (c->K < 2 || c->q[1] >= (1ull << 43))) INVALID("kind 5 needs modulus 1 of the chain below 2^43");
    TRY(set_device(c));
    TRY(arena_reserve(c, 256));
    const u32 iters = kind >= 3 ? 512 : 2048, grid = (u32)c->sms * 8, block = 256;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    float best = 0;
    u64 *o = (u64 *)c->arena.base;
    for (int rep = 0; rep < 4; ++rep) {  // first repetition warms up
        CU(cudaEventRecord(e0, c->stream));
        if (kind == 0) pipe_peak_kernel<0><<<grid, block, 0, c->stream>>>(o, iters, 12345u + rep, 0u, c->d_mods, c->d_modsd);
        if (kind == 1) pipe_peak_kernel<1><<<grid, block, 0, c->stream>>>(o, iters, 12345u + rep, 0u, c->d_mods, c->d_modsd);
        if (kind == 2) pipe_peak_kernel<2><<<grid, block, 0, c->stream>>>(o, iters, 12345u + rep, 0u, c->d_mods, c->d_modsd);
        if (kind == 3) pipe_peak_kernel<3><<<grid, block, 0, c->stream>>>(o, iters, 12345u + rep, 0u, c->d_mods, c->d_modsd);
        if (kind == 4) pipe_peak_kernel<4><<<grid, block, 0, c->stream>>>(o, iters, 12345u + rep, 0u, c->d_mods, c->d_modsd);
        if (kind == 5) pipe_peak_kernel<5><<<grid, block, 0, c->stream>>>(o, iters, 12345u + rep, 0u, c->d_mods, c->d_modsd);
        c->launches++;
        CU(cudaEventRecord(e1, c->stream));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (rep && (best == 0 || ms < best)) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ops_per_second = (double)grid * block * iters * 32.0 / (best * 1e-3);
    return HEGPU_OK;
}

// ------------------------------------------------------------------------- keys
static size_t key_words(hegpu_ctx *c) { return (size_t)(c->K - 1) * 2 * c->K * c->n; }
// device copies of key-switching keys are kept in Montgomery form (k * 2^64 mod m) for ks_inner
static int key_to_montgomery(hegpu_ctx *c, u64 *key)
{
    const size_t rows = (size_t)(c->K - 1) * 2 * c->K;
    to_montgomery_kernel<<<ew_grid(c, rows * c->n), 256, 0, c->stream>>>(key, key, rows, c->K, c->n, c->n, c->n, c->d_mods);
    c->launches++;
    CU(cudaGetLastError());
    return HEGPU_OK;
}

extern "C" int hegpu_load_relin_key(hegpu_ctx *c, const uint64_t *host)
{
    if (!c || !host) INVALID("null argument");
    if (c->K < 2) LOGIC("keyswitching is not supported by the context");
    TRY(set_device(c));
    if (!c->relin_key) CU(cudaMalloc(&c->relin_key, key_words(c) * sizeof(u64)));
    CU(cudaMemcpyAsync(c->relin_key, host, key_words(c) * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    TRY(key_to_montgomery(c, c->relin_key));
    CU(cudaStreamSynchronize(c->stream));
    return HEGPU_OK;
}

// GaloisTool::generate_table_ntt
static int get_perm(hegpu_ctx *c, u32 elt, const u32 **out)
{
    auto it = c->perms.find(elt);
    if (it != c->perms.end()) { *out = it->second; return HEGPU_OK; }
    std::vector<u32> tab(c->n);
    for (u32 i = 0; i < c->n; ++i) {
        u64 r = 2ull * h_brev(i, c->logn) + 1;
        u64 raw = ((u64)elt * r) >> 1;
        tab[i] = h_brev((u32)(raw & (c->n - 1)), c->logn);
    }
    u32 *d = nullptr;
    CU(cudaMalloc(&d, c->n * sizeof(u32)));
    CU(cudaMemcpy(d, tab.data(), c->n * sizeof(u32), cudaMemcpyHostToDevice));
    c->perms[elt] = d;
    *out = d;
    return HEGPU_OK;
}

extern "C" int hegpu_load_galois_key(hegpu_ctx *c, uint32_t elt, const uint64_t *host)
{
    if (!c || !host) INVALID("null argument");
    if (c->K < 2) LOGIC("keyswitching is not supported by the context");
    if (!(elt & 1) || elt >= 2 * c->n) INVALID("Galois element is not valid");
    TRY(set_device(c));
    c->key_epoch++;
    u64 *&slot = c->galois_keys[elt];
    if (!slot) CU(cudaMalloc(&slot, key_words(c) * sizeof(u64)));
    CU(cudaMemcpyAsync(slot, host, key_words(c) * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    TRY(key_to_montgomery(c, slot));
    CU(cudaStreamSynchronize(c->stream));
    const u32 *pm;
    TRY(get_perm(c, elt, &pm));
    return HEGPU_OK;
}
extern "C" int hegpu_has_galois_key(hegpu_ctx *c, uint32_t elt) { return c && c->galois_keys.count(elt) ? 1 : 0; }

extern "C" int hegpu_galois_elt_from_step(hegpu_ctx *c, int step, uint32_t *elt)
{
    if (!c || !elt) INVALID("null argument");
    const u32 n = c->n, m = 2 * n;
    if (step == 0) { *elt = m - 1; return HEGPU_OK; }
    const u32 pos = (u32)std::abs(step);
    if (pos >= (n >> 1)) INVALID("step count too large");
    const u32 s = step < 0 ? (n >> 1) - pos : pos;
    u64 e = 1;
    for (u32 i = 0; i < s; ++i) e = (e * 3) & (m - 1);
    *elt = (u32)e;
    return HEGPU_OK;
}

// ------------------------------------------------------------------------- handles
extern "C" int hegpu_ct_create(hegpu_ctx *c, hegpu_ct **out, uint32_t batch, uint32_t size_cap, uint32_t L_cap)
{
    if (!c || !out) INVALID("null argument");
    if (batch == 0 || size_cap < 2 || size_cap > 3 || L_cap == 0 || L_cap > c->K) INVALID("invalid ciphertext batch shape");
    TRY(set_device(c));
    hegpu_ct *t = new hegpu_ct{ c, nullptr, batch, size_cap, L_cap, 0, 0, 1.0 };
    cudaError_t e = cudaMalloc(&t->d, (size_t)batch * size_cap * L_cap * c->n * sizeof(u64));
    if (e != cudaSuccess) {
        delete t;
        return fail(e == cudaErrorMemoryAllocation ? HEGPU_ERR_OUT_OF_MEMORY : HEGPU_ERR_CUDA, cudaGetErrorString(e));
    }
    *out = t;
    return HEGPU_OK;
}
extern "C" int hegpu_ct_destroy(hegpu_ct *t)
{
    if (!t) return HEGPU_OK;
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    if (t->ev_copy) {
        cudaEventSynchronize(t->ev_copy);
        cudaEventDestroy(t->ev_copy);
    }
    cudaFree(t->d);
    delete t;
    return HEGPU_OK;
}

// compute enqueued after an asynchronous copy of `t` waits for that copy
static int await_copy(const hegpu_ct *t)
{
    if (t && t->copy_pending) {
        CU(cudaStreamWaitEvent(t->ctx->stream, t->ev_copy, 0));
        t->copy_pending = false;
    }
    return HEGPU_OK;
}
// Asynchronous copies run on dedicated copy streams: they start after all compute enqueued
// before the call and overlap compute enqueued after it.  Needs the packed layout
// (size == size_cap, L == L_cap) and pinned host memory to be truly asynchronous.
static int ct_copy_async(hegpu_ct *t, u64 *host, bool upload)
{
    hegpu_ctx *c = t->ctx;
    TRY(set_device(c));
    if (t->size != t->size_cap || t->L != t->L_cap) INVALID("asynchronous copies need size == size_cap and L == L_cap");
    if (!t->ev_copy) CU(cudaEventCreateWithFlags(&t->ev_copy, cudaEventDisableTiming));
    cudaStream_t cs = upload ? c->copy_h2d : c->copy_d2h;
    if (t->copy_pending) CU(cudaStreamWaitEvent(cs, t->ev_copy, 0));  // order after an earlier async copy
    CU(cudaEventRecord(c->ev_fence, c->stream));
    CU(cudaStreamWaitEvent(cs, c->ev_fence, 0));
    const size_t bytes = (size_t)t->batch * t->size * t->L * c->n * sizeof(u64);
    if (upload)
        CU(cudaMemcpyAsync(t->d, host, bytes, cudaMemcpyHostToDevice, cs));
    else
        CU(cudaMemcpyAsync(host, t->d, bytes, cudaMemcpyDeviceToHost, cs));
    CU(cudaEventRecord(t->ev_copy, cs));
    t->copy_pending = true;
    return HEGPU_OK;
}
extern "C" int hegpu_ct_upload_async(hegpu_ct *t, const uint64_t *host, uint32_t size, uint32_t L, double scale)
{
    if (!t || !host) INVALID("null argument");
    if (size < 2 || size > t->size_cap || L == 0 || L > t->L_cap) INVALID("ciphertext does not fit the batch capacity");
    if (size != t->size_cap || L != t->L_cap) INVALID("asynchronous copies need size == size_cap and L == L_cap");
    t->size = size;
    t->L = L;
    t->scale = scale;
    return ct_copy_async(t, const_cast<u64 *>((const u64 *)host), true);
}
extern "C" int hegpu_ct_download_async(hegpu_ct *t, uint64_t *host)
{
    if (!t || !host) INVALID("null argument");
    if (!t->L) INVALID("ciphertext batch is uninitialised");
    return ct_copy_async(t, (u64 *)host, false);
}
extern "C" int hegpu_ct_copy_wait(hegpu_ct *t)
{
    if (!t) INVALID("null argument");
    if (t->ev_copy) CU(cudaEventSynchronize(t->ev_copy));
    return HEGPU_OK;
}

static int launch_repack(hegpu_ctx *c, CtView dst, CtView src, u32 B, u32 polys, u32 L)
{
    const size_t total = (size_t)B * polys * L * c->n;
    Prof pf(c, PK_ELEMENTWISE, total, total * 16);
    repack_kernel<<<ew_grid(c, total), 256, 0, c->stream>>>(dst, src, B, polys, L, c->n);
    c->launches++;
    CU(cudaGetLastError());
    return HEGPU_OK;
}

static int ct_io(hegpu_ct *t, u32 b0, u32 nb, u64 *host, bool upload)
{
    hegpu_ctx *c = t->ctx;
    TRY(set_device(c));
    TRY(await_copy(t));
    const size_t words = (size_t)nb * t->size * t->L * c->n;
    const bool packed = (t->size == t->size_cap && t->L == t->L_cap);
    CtView dv = t->view_at(b0);
    if (packed) {
        if (upload)
            CU(cudaMemcpyAsync(dv.p, host, words * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
        else
            CU(cudaMemcpyAsync(host, dv.p, words * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    } else {
        TRY(stage_reserve(c, words));
        CtView sv{ c->stage, (size_t)t->size * t->L * c->n, (size_t)t->L * c->n, (size_t)c->n };
        if (upload) {
            CU(cudaMemcpyAsync(c->stage, host, words * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
            TRY(launch_repack(c, dv, sv, nb, t->size, t->L));
        } else {
            TRY(launch_repack(c, sv, dv, nb, t->size, t->L));
            CU(cudaMemcpyAsync(host, c->stage, words * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
        }
    }
    if (!upload) CU(cudaStreamSynchronize(c->stream));
    return HEGPU_OK;
}

extern "C" int hegpu_ct_upload(hegpu_ct *t, const uint64_t *host, uint32_t size, uint32_t L, double scale)
{
    if (!t || !host) INVALID("null argument");
    if (size < 2 || size > t->size_cap || L == 0 || L > t->L_cap) INVALID("ciphertext does not fit the batch capacity");
    t->size = size;
    t->L = L;
    t->scale = scale;
    return ct_io(t, 0, t->batch, const_cast<u64 *>((const u64 *)host), true);
}
extern "C" int hegpu_ct_download(hegpu_ct *t, uint64_t *host)
{
    if (!t || !host) INVALID("null argument");
    if (!t->L) INVALID("ciphertext batch is uninitialised");
    return ct_io(t, 0, t->batch, (u64 *)host, false);
}
extern "C" int hegpu_ct_upload_one(hegpu_ct *t, uint32_t index, const uint64_t *host)
{
    if (!t || !host) INVALID("null argument");
    if (index >= t->batch || !t->L) INVALID("invalid batch index or uninitialised metadata");
    return ct_io(t, index, 1, const_cast<u64 *>((const u64 *)host), true);
}
extern "C" int hegpu_ct_download_one(hegpu_ct *t, uint32_t index, uint64_t *host)
{
    if (!t || !host) INVALID("null argument");
    if (index >= t->batch || !t->L) INVALID("invalid batch index or uninitialised metadata");
    return ct_io(t, index, 1, (u64 *)host, false);
}
extern "C" int hegpu_ct_info(const hegpu_ct *t, uint32_t *batch, uint32_t *size, uint32_t *L, double *scale)
{
    if (!t) INVALID("null argument");
    if (batch) *batch = t->batch;
    if (size) *size = t->size;
    if (L) *L = t->L;
    if (scale) *scale = t->scale;
    return HEGPU_OK;
}
extern "C" int hegpu_ct_set_scale(hegpu_ct *t, double scale)
{
    if (!t) INVALID("null argument");
    t->scale = scale;
    return HEGPU_OK;
}
extern "C" int hegpu_ct_set_meta(hegpu_ct *t, uint32_t size, uint32_t L, double scale)
{
    if (!t) INVALID("null argument");
    if (size < 2 || size > t->size_cap || L == 0 || L > t->L_cap) INVALID("size or level exceeds the batch's capacity");
    t->size = size;
    t->L = L;
    t->scale = scale;
    return HEGPU_OK;
}
extern "C" int hegpu_ct_device_view(hegpu_ct *t, void **dptr, size_t *sb, size_t *sp, size_t *sl)
{
    if (!t) INVALID("null argument");
    CtView v = t->view();
    if (dptr) *dptr = v.p;
    if (sb) *sb = v.sb;
    if (sp) *sp = v.sp;
    if (sl) *sl = v.sl;
    return HEGPU_OK;
}

template <int OP>
static int launch_ew(hegpu_ctx *c, CtView out, CtView a, CtView b, u32 B, u32 polys, u32 L)
{
    EwParams P{ out, a, b, B, polys, L, c->n };
    const size_t total = (size_t)B * polys * L * c->n;
    Prof pf(c, PK_ELEMENTWISE, total, total * ((OP == EW_NEG || OP == EW_COPY || OP == EW_NEGCOPY_B) ? 16 : 24));
    ew_kernel<OP><<<ew_grid(c, total), 256, 0, c->stream>>>(P, c->d_mods);
    c->launches++;
    CU(cudaGetLastError());
    return HEGPU_OK;
}

static int fits(const hegpu_ct *out, u32 batch, u32 size, u32 L)
{
    if (out->batch != batch) INVALID("destination batch size mismatch");
    if (size > out->size_cap || L > out->L_cap) INVALID("destination capacity too small");
    return await_copy(out);
}

extern "C" int hegpu_ct_copy(hegpu_ctx *c, hegpu_ct *dst, const hegpu_ct *src)
{
    if (!c || !dst || !src) INVALID("null argument");
    if (!src->L) INVALID("ciphertext batch is uninitialised");
    TRY(set_device(c));
    TRY(fits(dst, src->batch, src->size, src->L));
    if (dst != src) TRY(launch_ew<EW_COPY>(c, dst->view(), src->view(), src->view(), src->batch, src->size, src->L));
    dst->size = src->size;
    dst->L = src->L;
    dst->scale = src->scale;
    return HEGPU_OK;
}
extern "C" int hegpu_ct_copy_one(hegpu_ctx *c, hegpu_ct *dst, uint32_t di, const hegpu_ct *src, uint32_t si)
{
    if (!c || !dst || !src) INVALID("null argument");
    if (!src->L || di >= dst->batch || si >= src->batch) INVALID("invalid batch index");
    if (src->size > dst->size_cap || src->L > dst->L_cap) INVALID("destination capacity too small");
    if (dst->L && (dst->L != src->L || dst->size != src->size)) INVALID("metadata mismatch");
    TRY(set_device(c));
    TRY(await_copy(src));
    TRY(await_copy(dst));
    TRY(launch_ew<EW_COPY>(c, dst->view_at(di), src->view_at(si), src->view_at(si), 1, src->size, src->L));
    dst->size = src->size;
    dst->L = src->L;
    dst->scale = src->scale;
    return HEGPU_OK;
}

extern "C" int hegpu_pt_create(hegpu_ctx *c, hegpu_pt **out, uint32_t count, uint32_t L_cap)
{
    if (!c || !out) INVALID("null argument");
    if (count == 0 || L_cap == 0 || L_cap > c->K) INVALID("invalid plaintext set shape");
    TRY(set_device(c));
    hegpu_pt *t = new hegpu_pt{ c, nullptr, count, L_cap, 0, 1.0 };
    cudaError_t e = cudaMalloc(&t->d, (size_t)count * L_cap * c->n * sizeof(u64));
    if (e != cudaSuccess) {
        delete t;
        return fail(e == cudaErrorMemoryAllocation ? HEGPU_ERR_OUT_OF_MEMORY : HEGPU_ERR_CUDA, cudaGetErrorString(e));
    }
    *out = t;
    return HEGPU_OK;
}
extern "C" int hegpu_pt_destroy(hegpu_pt *t)
{
    if (!t) return HEGPU_OK;
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    cudaFree(t->d);
    cudaFree(t->d_mont);
    cudaFree(t->d_w);
    delete t;
    return HEGPU_OK;
}
static int pt_montgomery(hegpu_pt *t, const u64 **out)
{
    hegpu_ctx *c = t->ctx;
    if (!t->d_mont) CU(cudaMalloc(&t->d_mont, (size_t)t->count * t->L_cap * c->n * sizeof(u64)));
    if (!t->mont_valid) {
        const size_t rows = (size_t)t->count * t->L_cap;
        to_montgomery_kernel<<<ew_grid(c, rows * c->n), 256, 0, c->stream>>>(t->d, t->d_mont, rows, t->L_cap, c->n, c->n, c->n, c->d_mods,
                                                                             t->ext ? t->L : ~0u, c->K);
        c->launches++;
        CU(cudaGetLastError());
        t->mont_valid = true;
    }
    *out = t->d_mont;
    return HEGPU_OK;
}
static int pt_io(hegpu_pt *t, u32 i0, u32 cnt, u64 *host, bool upload)
{
    hegpu_ctx *c = t->ctx;
    TRY(set_device(c));
    if (upload) {
        t->mont_valid = false;
        t->w_key.clear();
    }
    const size_t row = (size_t)(t->L + (t->ext ? 1 : 0)) * c->n * sizeof(u64);
    u64 *d = t->d + i0 * t->stride();
    if (upload)
        CU(cudaMemcpy2DAsync(d, t->stride() * sizeof(u64), host, row, row, cnt, cudaMemcpyHostToDevice, c->stream));
    else {
        CU(cudaMemcpy2DAsync(host, row, d, t->stride() * sizeof(u64), row, cnt, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return HEGPU_OK;
}
extern "C" int hegpu_pt_upload(hegpu_pt *t, const uint64_t *host, uint32_t L, double scale)
{
    if (!t || !host) INVALID("null argument");
    if (L == 0 || L > t->L_cap) INVALID("plaintext does not fit the set capacity");
    t->L = L;
    t->scale = scale;
    t->ext = false;
    return pt_io(t, 0, t->count, const_cast<u64 *>((const u64 *)host), true);
}
extern "C" int hegpu_pt_upload_ext(hegpu_pt *t, const uint64_t *host, uint32_t L, double scale)
{
    if (!t || !host) INVALID("null argument");
    if (L == 0 || L + 1 > t->L_cap || L + 1 > t->ctx->K) INVALID("plaintext does not fit the set capacity");
    t->L = L;
    t->scale = scale;
    t->ext = true;
    return pt_io(t, 0, t->count, const_cast<u64 *>((const u64 *)host), true);
}
extern "C" int hegpu_pt_upload_one(hegpu_pt *t, uint32_t index, const uint64_t *host)
{
    if (!t || !host) INVALID("null argument");
    if (index >= t->count || !t->L) INVALID("invalid plaintext index or uninitialised metadata");
    return pt_io(t, index, 1, const_cast<u64 *>((const u64 *)host), true);
}
extern "C" int hegpu_pt_download_one(hegpu_pt *t, uint32_t index, uint64_t *host)
{
    if (!t || !host) INVALID("null argument");
    if (index >= t->count || !t->L) INVALID("invalid plaintext index or uninitialised metadata");
    return pt_io(t, index, 1, (u64 *)host, false);
}

static size_t inv_scratch_words(hegpu_ctx *c, size_t jobs) { return c->logn == 15 ? jobs * c->n : 0; }

static int ntt_device(hegpu_ctx *c, void *d, u32 count, u32 first_mod, u32 n_mods, bool inverse)
{
    if (!c || !d) INVALID("null argument");
    if (n_mods == 0 || first_mod + n_mods > c->K) INVALID("modulus index out of range");
    TRY(set_device(c));
    PlainJob job{ (const u64 *)d, (u64 *)d, first_mod, n_mods, c->n };
    if (!inverse) return launch_ntt_fwd(c, job, count, PK_NTT_FWD_PLAIN, 2);
    const size_t sw = inv_scratch_words(c, count);
    TRY(arena_reserve(c, align256(sw)));
    return launch_ntt_inv(c, job, count, (u64 *)c->arena.base, PK_NTT_INV_PLAIN);
}
extern "C" int hegpu_ntt_forward_device(hegpu_ctx *c, void *d, uint32_t count, uint32_t first_mod, uint32_t n_mods)
{
    return ntt_device(c, d, count, first_mod, n_mods, false);
}
extern "C" int hegpu_ntt_inverse_device(hegpu_ctx *c, void *d, uint32_t count, uint32_t first_mod, uint32_t n_mods)
{
    return ntt_device(c, d, count, first_mod, n_mods, true);
}
static int ntt_host(hegpu_ctx *c, u64 *host, u32 count, u32 first_mod, u32 n_mods, bool inverse)
{
    if (!c || !host) INVALID("null argument");
    TRY(set_device(c));
    const size_t words = (size_t)count * c->n;
    TRY(stage_reserve(c, words));
    CU(cudaMemcpyAsync(c->stage, host, words * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    TRY(ntt_device(c, c->stage, count, first_mod, n_mods, inverse));
    CU(cudaMemcpyAsync(host, c->stage, words * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return HEGPU_OK;
}
extern "C" int hegpu_ntt_forward_host(hegpu_ctx *c, uint64_t *h, uint32_t count, uint32_t first_mod, uint32_t n_mods)
{
    return ntt_host(c, (u64 *)h, count, first_mod, n_mods, false);
}
extern "C" int hegpu_ntt_inverse_host(hegpu_ctx *c, uint64_t *h, uint32_t count, uint32_t first_mod, uint32_t n_mods)
{
    return ntt_host(c, (u64 *)h, count, first_mod, n_mods, true);
}

// ------------------------------------------------------------------------- checks (SEAL wording)
static bool are_close(double a, double b)
{
    double sf = std::max(std::max(std::fabs(a), std::fabs(b)), 1.0);
    return std::fabs(a - b) < std::numeric_limits<double>::epsilon() * sf;
}
static bool scale_in_bounds(hegpu_ctx *c, double scale, u32 L)
{
    if (!(scale > 0)) return false;
    return (int)std::log2(scale) < c->level_bits[L - 1];
}
static int check_ct(const hegpu_ct *a)
{
    if (!a) INVALID("null argument");
    if (!a->L || a->size < 2) INVALID("encrypted is not valid for encryption parameters");
    return await_copy(a);
}
static int batch_of(const hegpu_ct *a, const hegpu_ct *b, u32 *B, size_t *b_sb)
{
    if (b->batch != a->batch && b->batch != 1) INVALID("batch sizes do not match");
    *B = a->batch;
    *b_sb = (b->batch == 1 && a->batch != 1) ? 0 : b->view().sb;
    return HEGPU_OK;
}

// ------------------------------------------------------------------------- element-wise ops
extern "C" int hegpu_negate(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *a)
{
    if (!c || !out) INVALID("null argument");
    TRY(check_ct(a));
    TRY(set_device(c));
    TRY(fits(out, a->batch, a->size, a->L));
    TRY(launch_ew<EW_NEG>(c, out->view(), a->view(), a->view(), a->batch, a->size, a->L));
    out->size = a->size;
    out->L = a->L;
    out->scale = a->scale;
    return HEGPU_OK;
}

static int add_sub(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *a, const hegpu_ct *b, bool sub)
{
    if (!c || !out) INVALID("null argument");
    TRY(check_ct(a));
    TRY(check_ct(b));
    if (a->L != b->L) INVALID("encrypted1 and encrypted2 parameter mismatch");
    if (!are_close(a->scale, b->scale)) INVALID("scale mismatch");
    TRY(set_device(c));
    u32 B;
    size_t bsb;
    TRY(batch_of(a, b, &B, &bsb));
    const u32 mn = std::min(a->size, b->size), mx = std::max(a->size, b->size);
    TRY(fits(out, B, mx, a->L));
    CtView va = a->view(), vb = b->view(), vo = out->view();
    vb.sb = bsb;
    if (sub)
        TRY(launch_ew<EW_SUB>(c, vo, va, vb, B, mn, a->L));
    else
        TRY(launch_ew<EW_ADD>(c, vo, va, vb, B, mn, a->L));
    if (mx > mn) {  // extra polynomial of the larger operand: copied (add / minuend) or negated (subtrahend)
        CtView xo = vo, xa = va, xb = vb;
        xo.p += mn * xo.sp;
        xa.p += mn * xa.sp;
        xb.p += mn * xb.sp;
        if (a->size > b->size) {
            if (out != a) TRY(launch_ew<EW_COPY>(c, xo, xa, xa, B, mx - mn, a->L));
        } else if (sub) {
            TRY(launch_ew<EW_NEGCOPY_B>(c, xo, xb, xb, B, mx - mn, a->L));
        } else {
            TRY(launch_ew<EW_COPY>(c, xo, xb, xb, B, mx - mn, a->L));
        }
    }
    out->size = mx;
    out->L = a->L;
    out->scale = a->scale;
    return HEGPU_OK;
}
extern "C" int hegpu_add(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *a, const hegpu_ct *b) { return add_sub(c, out, a, b, false); }
extern "C" int hegpu_sub(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *a, const hegpu_ct *b) { return add_sub(c, out, a, b, true); }

static int pt_view(const hegpu_ct *a, const hegpu_pt *pt, int pt_index, CtView *v)
{
    if (!pt || !pt->L) INVALID("plain is not valid for encryption parameters");
    if (pt->L != a->L) INVALID("encrypted_ntt and plain_ntt parameter mismatch");
    if (pt_index >= 0) {
        if ((u32)pt_index >= pt->count) INVALID("plaintext index out of range");
        *v = CtView{ pt->d + (size_t)pt_index * pt->stride(), 0, 0, (size_t)pt->ctx->n };
    } else {
        if (pt->count != a->batch) INVALID("plaintext set size must equal the batch size");
        *v = CtView{ pt->d, pt->stride(), 0, (size_t)pt->ctx->n };
    }
    return HEGPU_OK;
}

static int plain_addsub(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *a, const hegpu_pt *pt, int pt_index, bool sub)
{
    if (!c || !out) INVALID("null argument");
    TRY(check_ct(a));
    CtView vp;
    TRY(pt_view(a, pt, pt_index, &vp));
    if (!are_close(a->scale, pt->scale)) INVALID("scale mismatch");
    TRY(set_device(c));
    TRY(fits(out, a->batch, a->size, a->L));
    CtView va = a->view(), vo = out->view();
    if (sub)
        TRY(launch_ew<EW_SUBPLAIN>(c, vo, va, vp, a->batch, 1, a->L));
    else
        TRY(launch_ew<EW_ADDPLAIN>(c, vo, va, vp, a->batch, 1, a->L));
    if (out != a) {
        CtView xo = vo, xa = va;
        xo.p += xo.sp;
        xa.p += xa.sp;
        TRY(launch_ew<EW_COPY>(c, xo, xa, xa, a->batch, a->size - 1, a->L));
    }
    out->size = a->size;
    out->L = a->L;
    out->scale = a->scale;
    return HEGPU_OK;
}
extern "C" int hegpu_add_plain(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *a, const hegpu_pt *pt, int i)
{
    return plain_addsub(c, out, a, pt, i, false);
}
extern "C" int hegpu_sub_plain(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *a, const hegpu_pt *pt, int i)
{
    return plain_addsub(c, out, a, pt, i, true);
}

extern "C" int hegpu_multiply_plain(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *a, const hegpu_pt *pt, int pt_index)
{
    if (!c || !out) INVALID("null argument");
    TRY(check_ct(a));
    CtView vp;
    TRY(pt_view(a, pt, pt_index, &vp));
    const double ns = a->scale * pt->scale;
    if (!scale_in_bounds(c, ns, a->L)) INVALID("scale out of bounds");
    TRY(set_device(c));
    TRY(fits(out, a->batch, a->size, a->L));
    TRY(launch_ew<EW_MULPLAIN>(c, out->view(), a->view(), vp, a->batch, a->size, a->L));
    out->size = a->size;
    out->L = a->L;
    out->scale = ns;
    return HEGPU_OK;
}

static int multiply_impl(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *a, const hegpu_ct *b, bool square)
{
    if (!c || !out) INVALID("null argument");
    TRY(check_ct(a));
    TRY(check_ct(b));
    if (a->L != b->L) INVALID("encrypted1 and encrypted2 parameter mismatch");
    const double ns = a->scale * b->scale;
    if (!scale_in_bounds(c, ns, a->L)) INVALID("scale out of bounds");
    TRY(set_device(c));
    u32 B;
    size_t bsb;
    TRY(batch_of(a, b, &B, &bsb));
    const u32 so = a->size + b->size - 1;
    if (so > 3) INVALID("ciphertext products beyond size 3 are not supported; relinearize first");
    TRY(fits(out, B, so, a->L));
    CtView vb = b->view();
    vb.sb = bsb;
    EwParams P{ out->view(), a->view(), vb, B, so, a->L, c->n };
    const size_t total = (size_t)B * a->L * c->n;
    Prof pf(c, PK_TENSOR, total, total * (square ? 40 : 56));
    if (a->size == 2 && b->size == 2) {
        if (square)
            tensor_kernel<true><<<ew_grid(c, total), 256, 0, c->stream>>>(P, c->d_mods);
        else
            tensor_kernel<false><<<ew_grid(c, total), 256, 0, c->stream>>>(P, c->d_mods);
    } else {
        INVALID("ciphertext products beyond size 3 are not supported; relinearize first");
    }
    c->launches++;
    CU(cudaGetLastError());
    out->size = so;
    out->L = a->L;
    out->scale = ns;
    return HEGPU_OK;
}
extern "C" int hegpu_multiply(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *a, const hegpu_ct *b) { return multiply_impl(c, out, a, b, false); }
extern "C" int hegpu_square(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *a) { return multiply_impl(c, out, a, a, true); }

// ------------------------------------------------------------------------- rescale / mod switch
// a: view of B ciphertexts with `size` polys at level L; out at level L-1 (may alias a)
static size_t rescale_scratch(hegpu_ctx *c, u32 B, u32 size) { return align256((size_t)B * size * c->n) + align256(inv_scratch_words(c, (size_t)B * size)); }
static int rescale_views(hegpu_ctx *c, CtView out, CtView a, u32 B, u32 size, u32 L, ArenaPlan &ap, bool lazy_in = false)
{
    const u32 jobs = B * size;
    u64 *t = ap.take((size_t)jobs * c->n);
    u64 *scr = ap.take(inv_scratch_words(c, jobs));
    HalfInttJob hj{ a.p + (size_t)(L - 1) * a.sl, t, a.sb, a.sp, size, L - 1, c->n, lazy_in ? 1u : 0u };
    TRY(launch_ntt_inv(c, hj, jobs, scr, PK_HALF_INTT));
    RescaleJob rj{ a, out, t, c->d_md + (size_t)(L - 1) * c->K, c->d_mods, size, L - 1, L - 1, c->n, lazy_in ? 1u : 0u };
    TRY(launch_ntt_fwd(c, rj, jobs * (L - 1), PK_RESCALE_NTT, 3));
    return HEGPU_OK;
}

extern "C" int hegpu_rescale_to_next(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *a)
{
    if (!c || !out) INVALID("null argument");
    TRY(check_ct(a));
    if (a->L < 2) INVALID("end of modulus switching chain reached");
    TRY(set_device(c));
    TRY(fits(out, a->batch, a->size, a->L - 1));
    TRY(arena_reserve(c, rescale_scratch(c, a->batch, a->size)));
    ArenaPlan ap{ c };
    TRY(rescale_views(c, out->view(), a->view(), a->batch, a->size, a->L, ap));
    out->size = a->size;
    out->scale = a->scale / (double)c->q[a->L - 1];
    out->L = a->L - 1;
    return HEGPU_OK;
}

// rescale of the element-wise uint64 SUM of up to `terms` partial ciphertexts (what an NCCL sum of the
// diagonal shards leaves behind): the reduction to [0,q) rides in the loads of the rescale's two transforms
extern "C" int hegpu_rescale_sum_to_next(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *a, uint32_t terms)
{
    if (!c || !out) INVALID("null argument");
    TRY(check_ct(a));
    if (terms == 0 || terms > 16) INVALID("at most 16 partial sums of 60-bit residues fit in 64 bits");
    if (a->L < 2) INVALID("end of modulus switching chain reached");
    if (out == a) INVALID("the rescale of a partial sum must not run in place");  // the epilogue re-reads the unreduced words
    TRY(set_device(c));
    TRY(fits(out, a->batch, a->size, a->L - 1));
    TRY(arena_reserve(c, rescale_scratch(c, a->batch, a->size)));
    ArenaPlan ap{ c };
    TRY(rescale_views(c, out->view(), a->view(), a->batch, a->size, a->L, ap, true));
    out->size = a->size;
    out->scale = a->scale / (double)c->q[a->L - 1];
    out->L = a->L - 1;
    return HEGPU_OK;
}

extern "C" int hegpu_mod_switch_to_next(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *a)
{
    if (!c || !out) INVALID("null argument");
    TRY(check_ct(a));
    if (a->L < 2) INVALID("end of modulus switching chain reached");
    if (!scale_in_bounds(c, a->scale, a->L - 1)) INVALID("scale out of bounds");  // SEAL mod_switch_drop_to_next
    TRY(set_device(c));
    TRY(fits(out, a->batch, a->size, a->L - 1));
    if (out != a) TRY(launch_ew<EW_COPY>(c, out->view(), a->view(), a->view(), a->batch, a->size, a->L - 1));
    out->size = a->size;
    out->scale = a->scale;
    out->L = a->L - 1;
    return HEGPU_OK;
}

// ------------------------------------------------------------------------- key switching
struct KsGroupDesc {
    CtView in, out;
    const u64 *key;
    const u32 *perm;
};
static size_t ks_scratch(hegpu_ctx *c, size_t E, u32 L)
{
    const size_t n = c->n;
    return align256(E * L * n) + align256(E * L * (L + 1) * n) + align256(E * 2 * (L + 1) * n) + align256(E * 2 * n) +
           align256(inv_scratch_words(c, E * std::max<u32>(L, 2)));
}
// Evaluator::switch_key_inplace for ngroups x B ciphertexts, in three phases so
// that composites can hoist the decomposition or defer the mod-down.
struct KsPlan {
    KsParams P;
    size_t E;   // ngroups * B
    u64 *scr;   // INTT scratch (N = 32768 only)
};
static int ks_setup(hegpu_ctx *c, KsPlan &pl, const KsGroupDesc *groups, u32 ngroups, u32 B, u32 L, u32 target_poly,
                    bool has_base1, bool hoisted, ArenaPlan &ap)
{
    const size_t E = (size_t)ngroups * B, n = c->n;
    const size_t EA = hoisted ? B : E;  // elements that are decomposed
    KsParams &P = pl.P;
    P = KsParams{};
    P.ngroups = ngroups;
    P.B = B;
    P.L = L;
    P.K = c->K;
    P.n = c->n;
    P.target_poly = target_poly;
    P.has_base1 = has_base1 ? 1u : 0u;
    P.hoisted = hoisted ? 1u : 0u;
    for (u32 g = 0; g < ngroups; ++g) {
        P.in[g] = groups[g].in;
        P.out[g] = groups[g].out;
        P.key[g] = groups[g].key;
        P.perm[g] = groups[g].perm;
    }
    P.coef = ap.take(EA * L * n);
    P.ext = ap.take(EA * L * (L + 1) * n);
    P.acc = ap.take(E * 2 * (L + 1) * n);
    P.t = ap.take(E * 2 * n);
    pl.E = E;
    pl.scr = ap.take(inv_scratch_words(c, E * std::max<u32>(L, 2)));
    return HEGPU_OK;
}
// steps 1-2: c_j = INTT(pi(target)_j); ext[j][i] = NTT_{m_i}(c_j mod m_i)
static int ks_decompose(hegpu_ctx *c, KsPlan &pl)
{
    const u32 L = pl.P.L;
    const size_t EA = pl.P.hoisted ? pl.P.B : pl.E;
    KsParams PA = pl.P;
    if (pl.P.hoisted) PA.ngroups = 1;
    KsInttJob j1{ PA };
    TRY(launch_ntt_inv(c, j1, (u32)(EA * L), pl.scr, PK_KS_INTT));
    KsLiftJob j2{ PA, c->d_mods };
    TRY(launch_ntt_fwd(c, j2, (u32)(EA * L * L), PK_KS_LIFT_NTT, 2));
    return HEGPU_OK;
}
// step 3: key inner product into acc[E][2][L+1][N]
static int ks_inner(hegpu_ctx *c, KsPlan &pl)
{
    const KsParams &P = pl.P;
    const u32 L = P.L, B = P.B, ngroups = P.ngroups;
    const size_t total = pl.E * (L + 1) * c->n;
    // per (e,i,x): L digits read + 2 written; the 2L key words are read once per batch chunk
    Prof pf(c, PK_KS_INNER, total, total * 8 * (L + 2) + (u64)ngroups * (L + 1) * c->n * 16 * L);
    dim3 grid((c->n + 255) / 256, ngroups * (L + 1), (B + KS_INNER_BCHUNK - 1) / KS_INNER_BCHUNK);
    switch (L) {
    case 1: ks_inner_kernel<1><<<grid, 256, 0, c->stream>>>(P, c->d_mods); break;
    case 2: ks_inner_kernel<2><<<grid, 256, 0, c->stream>>>(P, c->d_mods); break;
    case 3: ks_inner_kernel<3><<<grid, 256, 0, c->stream>>>(P, c->d_mods); break;
    case 4: ks_inner_kernel<4><<<grid, 256, 0, c->stream>>>(P, c->d_mods); break;
    default: ks_inner_kernel<0><<<grid, 256, 0, c->stream>>>(P, c->d_mods); break;
    }
    c->launches++;
    CU(cudaGetLastError());
    return HEGPU_OK;
}
// step 3 for lazy giant steps: out = init + sum over the plan's groups of the key inner products
static int ks_inner_sum(hegpu_ctx *c, KsPlan &pl, const u64 *init, u64 *out)
{
    const KsParams &P = pl.P;
    const u32 L = P.L, B = P.B, ngroups = P.ngroups;
    if (L > 4 || P.hoisted) {  // generic shapes: one accumulator per group, then the group sum
        TRY(ks_inner(c, pl));
        const size_t total = (size_t)B * 2 * (L + 1) * c->n;
        Prof pf(c, PK_ELEMENTWISE, total, total * 8 * (ngroups + 1 + (init ? 1 : 0)));
        GatherU0 G{};
        G.u0 = P.u0;
        for (u32 g = 0; g < ngroups; ++g) G.perm[g] = P.perm[g];
        acc_group_sum_kernel<<<ew_grid(c, total), 256, 0, c->stream>>>(P.acc, init, out, ngroups, B, L, c->K, c->n, c->d_mods, G);
        c->launches++;
        CU(cudaGetLastError());
        return HEGPU_OK;
    }
    const size_t total = (size_t)B * (L + 1) * c->n;
    // per (b,i,x): L digits per group read, 2 written (+2 init); keys read once per batch chunk
    Prof pf(c, PK_KS_INNER, total * ngroups, total * 8 * ((u64)ngroups * L + 2 + (init ? 2 : 0)) + (u64)ngroups * (L + 1) * c->n * 16 * L);
    dim3 grid((c->n + 255) / 256, L + 1, (B + KS_INNER_BCHUNK - 1) / KS_INNER_BCHUNK);
    switch (L) {
    case 1: ks_inner_sum_kernel<1><<<grid, 256, 0, c->stream>>>(P, init, out, c->d_mods); break;
    case 2: ks_inner_sum_kernel<2><<<grid, 256, 0, c->stream>>>(P, init, out, c->d_mods); break;
    case 3: ks_inner_sum_kernel<3><<<grid, 256, 0, c->stream>>>(P, init, out, c->d_mods); break;
    default: ks_inner_sum_kernel<4><<<grid, 256, 0, c->stream>>>(P, init, out, c->d_mods); break;
    }
    c->launches++;
    CU(cudaGetLastError());
    return HEGPU_OK;
}
// steps 4-5: mod-down by P with rounding, add the base ciphertext, write out
static int ks_moddown(hegpu_ctx *c, KsPlan &pl)
{
    const KsParams &P = pl.P;
    const u32 L = P.L;
    if (P.only_c1) {  // component 1 of every element: t is [E][N]
        HalfInttJob j4{ P.acc + (size_t)(2 * L + 1) * c->n, P.t, (size_t)2 * (L + 1) * c->n, 0, 1, c->K - 1, c->n, 0 };
        TRY(launch_ntt_inv(c, j4, (u32)pl.E, pl.scr, PK_HALF_INTT));
        KsModDownJob j5{ P, c->d_md + (size_t)(c->K - 1) * c->K, c->d_mods, c->limb_major ? (u32)pl.E : 0u };
        TRY(launch_ntt_fwd(c, j5, (u32)(pl.E * L), PK_KS_MODDOWN_NTT, 3));
        return HEGPU_OK;
    }
    HalfInttJob j4{ P.acc + (size_t)L * c->n, P.t, (size_t)(L + 1) * c->n, 0, 1, c->K - 1, c->n, 0 };
    TRY(launch_ntt_inv(c, j4, (u32)(pl.E * 2), pl.scr, PK_HALF_INTT));
    KsModDownJob j5{ P, c->d_md + (size_t)(c->K - 1) * c->K, c->d_mods, c->limb_major ? (u32)(pl.E * 2) : 0u };
    TRY(launch_ntt_fwd(c, j5, (u32)(pl.E * 2 * L), PK_KS_MODDOWN_NTT, P.has_base1 ? 4 : 3));
    return HEGPU_OK;
}
// steps 4-5 and the rescale that follows, fused (FinalInttJob / FinalNttJob): out (level L-1) =
// rescale(base + mod_down(acc)), bit-identical to ks_moddown + rescale_views with 8 transforms per
// ciphertext instead of 14 (L = 3).  pl: ngroups = 1, acc, t, in[0] = base (has_base1 / no_base0), scr.
static int ks_moddown_rescale(hegpu_ctx *c, KsPlan &pl, CtView out, u64 *t2)
{
    const KsParams &P = pl.P;
    const u32 L = P.L, B = P.B;
    HalfInttJob j4{ P.acc + (size_t)L * c->n, P.t, (size_t)(L + 1) * c->n, 0, 1, c->K - 1, c->n, 0 };
    TRY(launch_ntt_inv(c, j4, B * 2, pl.scr, PK_HALF_INTT));
    FinalParams F{};
    F.acc = P.acc;
    F.base = P.in[0];
    F.out = out;
    F.t = P.t;
    F.t2 = t2;
    F.mdP = c->d_md + (size_t)(c->K - 1) * c->K;
    F.mdQ = c->d_md + (size_t)(L - 1) * c->K;
    F.mods = c->d_mods;
    F.B = B;
    F.L = L;
    F.K = c->K;
    F.n = c->n;
    F.has_base0 = P.no_base0 ? 0u : 1u;
    F.has_base1 = P.has_base1;
    F.per = c->limb_major ? B * 2 : 0u;
    FinalInttJob ji{ F };
    TRY(launch_ntt_inv(c, ji, B * 2, pl.scr, PK_HALF_INTT));
    FinalNttJob jn{ F };
    TRY(launch_ntt_fwd(c, jn, B * 2 * (L - 1), PK_RESCALE_NTT, 5));
    return HEGPU_OK;
}
// out must not alias in when perm != null.
static int keyswitch(hegpu_ctx *c, const KsGroupDesc *groups, u32 ngroups, u32 B, u32 L, u32 target_poly, bool has_base1,
                     ArenaPlan &ap, bool hoisted = false)
{
    if (ngroups == 0 || B == 0) return HEGPU_OK;
    KsPlan pl;
    TRY(ks_setup(c, pl, groups, ngroups, B, L, target_poly, has_base1, hoisted, ap));
    TRY(ks_decompose(c, pl));
    TRY(ks_inner(c, pl));
    TRY(ks_moddown(c, pl));
    return HEGPU_OK;
}

extern "C" int hegpu_relinearize(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *a)
{
    if (!c || !out) INVALID("null argument");
    TRY(check_ct(a));
    if (a->size == 2) return hegpu_ct_copy(c, out, a);
    if (a->size != 3) INVALID("encrypted size must be 2 or 3");
    if (!c->relin_key) INVALID("not enough relinearization keys");
    TRY(set_device(c));
    TRY(fits(out, a->batch, 2, a->L));
    TRY(arena_reserve(c, ks_scratch(c, a->batch, a->L)));
    ArenaPlan ap{ c };
    KsGroupDesc g{ a->view(), out->view(), c->relin_key, nullptr };
    TRY(keyswitch(c, &g, 1, a->batch, a->L, 2, true, ap));
    out->size = 2;
    out->L = a->L;
    out->scale = a->scale;
    return HEGPU_OK;
}

// out (view) = apply_galois(in (view)); views must not alias
static int galois_views(hegpu_ctx *c, CtView out, CtView in, u32 B, u32 L, u32 elt, ArenaPlan &ap)
{
    auto it = c->galois_keys.find(elt);
    if (it == c->galois_keys.end()) INVALID("Galois key not present");
    const u32 *pm;
    TRY(get_perm(c, elt, &pm));
    KsGroupDesc g{ in, out, it->second, pm };
    return keyswitch(c, &g, 1, B, L, 1, false, ap);
}

extern "C" int hegpu_apply_galois(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *a, uint32_t elt)
{
    if (!c || !out) INVALID("null argument");
    TRY(check_ct(a));
    if (a->size > 2) INVALID("encrypted size must be 2");
    if (!(elt & 1) || elt >= 2 * c->n) INVALID("Galois element is not valid");
    if (!c->galois_keys.count(elt)) INVALID("Galois key not present");
    TRY(set_device(c));
    TRY(fits(out, a->batch, 2, a->L));
    const size_t ctw = align256((size_t)a->batch * 2 * a->L * c->n);
    TRY(arena_reserve(c, ks_scratch(c, a->batch, a->L) + ctw));
    ArenaPlan ap{ c };
    if (out == a || out->d == a->d) {
        u64 *tmp = ap.take((size_t)a->batch * 2 * a->L * c->n);
        CtView tv{ tmp, (size_t)2 * a->L * c->n, (size_t)a->L * c->n, (size_t)c->n };
        TRY(galois_views(c, tv, a->view(), a->batch, a->L, elt, ap));
        TRY(launch_ew<EW_COPY>(c, out->view(), tv, tv, a->batch, 2, a->L));
    } else {
        TRY(galois_views(c, out->view(), a->view(), a->batch, a->L, elt, ap));
    }
    out->size = 2;
    out->L = a->L;
    out->scale = a->scale;
    return HEGPU_OK;
}

// util::naf
static std::vector<int> naf(int value)
{
    std::vector<int> res;
    bool sign = value < 0;
    value = std::abs(value);
    for (int i = 0; value; ++i) {
        int zi = (value & 1) ? 2 - (value & 3) : 0;
        value = (value - zi) >> 1;
        if (zi) res.push_back((sign ? -zi : zi) * (1 << i));
    }
    return res;
}

// Evaluator::rotate_internal, in place on `t`
static int rotate_internal(hegpu_ctx *c, hegpu_ct *t, int steps)
{
    if (steps == 0) return HEGPU_OK;
    u32 elt;
    TRY(hegpu_galois_elt_from_step(c, steps, &elt));
    if (c->galois_keys.count(elt)) return hegpu_apply_galois(c, t, t, elt);
    std::vector<int> terms = naf(steps);
    if (terms.size() == 1) INVALID("Galois key not present");
    for (int s : terms)
        if ((u32)std::abs(s) != (c->n >> 1)) TRY(rotate_internal(c, t, s));
    return HEGPU_OK;
}

extern "C" int hegpu_rotate_vector(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *a, int steps)
{
    if (!c || !out) INVALID("null argument");
    TRY(check_ct(a));
    if (a->size > 2) INVALID("encrypted size must be 2");
    u32 elt;
    TRY(hegpu_galois_elt_from_step(c, steps, &elt));
    if (c->galois_keys.empty() && steps != 0) INVALID("Galois key not present");
    if (steps != 0 && c->galois_keys.count(elt) && out != a) return hegpu_apply_galois(c, out, a, elt);
    TRY(hegpu_ct_copy(c, out, a));
    return rotate_internal(c, out, steps);
}

// ------------------------------------------------------------------------- multi-GPU fix-up
extern "C" int hegpu_reduce_fixup(hegpu_ctx *c, hegpu_ct *t, uint32_t terms)
{
    if (!c) INVALID("null argument");
    TRY(check_ct(t));
    if (terms == 0 || terms > 16) INVALID("at most 16 partial sums of 60-bit residues fit in 64 bits");
    TRY(set_device(c));
    const size_t total = (size_t)t->batch * t->size * t->L * c->n;
    Prof pf(c, PK_ELEMENTWISE, total, total * 16);
    fixup_kernel<<<ew_grid(c, total), 256, 0, c->stream>>>(t->view(), t->batch, t->size, t->L, c->n, c->d_mods);
    c->launches++;
    CU(cudaGetLastError());
    return HEGPU_OK;
}

// ------------------------------------------------------------------------- transparency query
extern "C" int hegpu_ct_transparent(hegpu_ctx *c, const hegpu_ct *a, uint32_t *count)
{
    if (!c || !count) INVALID("null argument");
    TRY(check_ct(a));
    TRY(set_device(c));
    *count = 0;
    if (a->batch == 0) return HEGPU_OK;
    TRY(arena_reserve(c, align256((a->batch + 1) / 2)));
    ArenaPlan ap{ c };
    u32 *flags = reinterpret_cast<u32 *>(ap.take((a->batch + 1) / 2));
    CU(cudaMemsetAsync(flags, 0, sizeof(u32) * a->batch, c->stream));
    const size_t total = (size_t)a->batch * (a->size - 1) * a->L * c->n;
    {
        Prof pf(c, PK_ELEMENTWISE, total, total * 8);
        nonzero_tail_kernel<<<ew_grid(c, total), 256, 0, c->stream>>>(a->view(), a->batch, a->size, a->L, c->n, flags);
        c->launches++;
        CU(cudaGetLastError());
    }
    std::vector<u32> h(a->batch);
    CU(cudaMemcpyAsync(h.data(), flags, sizeof(u32) * a->batch, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (u32 f : h) *count += f ? 0u : 1u;
    return HEGPU_OK;
}

// ------------------------------------------------------------------------- composites
template <int N1>
static void launch_bsgs_inner(hegpu_ctx *c, const BsgsParams &P)
{
    dim3 grid(P.n / 32, P.L, (P.n2 + BSGS_GT - 1) / BSGS_GT), block(32, BSGS_GT);
    bsgs_inner_kernel<N1><<<grid, block, 0, c->stream>>>(P, c->d_mods);
}
static void launch_bsgs_inner_n1(hegpu_ctx *c, const BsgsParams &P, u32 n1)
{
    switch (n1) {
    case 1: launch_bsgs_inner<1>(c, P); break;
    case 2: launch_bsgs_inner<2>(c, P); break;
    case 3: launch_bsgs_inner<3>(c, P); break;
    case 4: launch_bsgs_inner<4>(c, P); break;
    case 5: launch_bsgs_inner<5>(c, P); break;
    case 6: launch_bsgs_inner<6>(c, P); break;
    case 7: launch_bsgs_inner<7>(c, P); break;
    case 8: launch_bsgs_inner<8>(c, P); break;
    case 9: launch_bsgs_inner<9>(c, P); break;
    case 10: launch_bsgs_inner<10>(c, P); break;
    case 11: launch_bsgs_inner<11>(c, P); break;
    case 12: launch_bsgs_inner<12>(c, P); break;
    case 13: launch_bsgs_inner<13>(c, P); break;
    case 14: launch_bsgs_inner<14>(c, P); break;
    case 15: launch_bsgs_inner<15>(c, P); break;
    case 16: launch_bsgs_inner<16>(c, P); break;
    case 24: launch_bsgs_inner<24>(c, P); break;
    default: launch_bsgs_inner<32>(c, P); break;
    }
}

static u32 dh_n2_pad(u32 n2) { return n2 <= 1 ? 1 : n2 <= 2 ? 2 : 4; }  // giant steps per launch
template <int LT, int N2>
static int launch_dh_inner(hegpu_ctx *c, const DhInnerParams &P)
{
    const size_t smem = dh_inner_smem(P.n1, N2, LT);
    auto kern = dh_inner_kernel<LT, N2>;
    DhInnerParams Q = P;
    Q.stream_out = (u32)c->dh_stcs;
    TRY(configure_smem(c, (const void *)kern, smem));
    dim3 grid(P.n / DH_TX * (LT + 1), 1, (P.B + DH_BCH - 1) / DH_BCH);
    const bool use_bk = P.bk != nullptr && N2 == 4 && P.n2 > (u32)N2;
    if (!use_bk) Q.bk = nullptr;
    if constexpr (N2 == 4) {
        if (use_bk) {
            TRY(configure_smem(c, (const void *)dh_inner_bk_kernel<LT, N2, 1>, smem));
            TRY(configure_smem(c, (const void *)dh_inner_bk_kernel<LT, N2, 2>, smem));
        }
    }
    for (u32 g0 = 0; g0 < P.n2; g0 += N2) {
        Q.g0 = g0;
        Q.ng = std::min<u32>(N2, P.n2 - g0);
        Q.bk_mode = !use_bk ? 0u : (g0 == 0 ? 1u : 2u);
        const int swz = (c->dh_swz && LT == 3 && c->sms % 4 == 0) ? (c->sms << 8) : 0;
        bool launched = false;
        if constexpr (N2 == 4) {
            if (Q.bk_mode == 1) {
                dh_inner_bk_kernel<LT, N2, 1><<<grid, DH_TX * DH_KG, smem, c->stream>>>(Q, c->d_mods, c->dh_f64 | swz);
                launched = true;
            } else if (Q.bk_mode == 2) {
                dh_inner_bk_kernel<LT, N2, 2><<<grid, DH_TX * DH_KG, smem, c->stream>>>(Q, c->d_mods, c->dh_f64 | swz);
                launched = true;
            }
        }
        if (!launched) kern<<<grid, DH_TX * DH_KG, smem, c->stream>>>(Q, c->d_mods, c->dh_f64 | swz);
        c->launches++;
        CU(cudaGetLastError());
    }
    return HEGPU_OK;
}

// HEGPU_MATVEC_IMMA: W = baby-step keys (.) diagonals as MMA B fragments, one block per launch group of 4 giant steps;
// rebuilt only when the diagonals, the shape or the key set change
static int pt_imma_w(hegpu_ctx *c, hegpu_pt *t, u32 n1, u32 n2, u32 g_first, u32 L, const std::vector<u32> &belt, const u32 **w, u32 *ksteps_out,
                     size_t *group_words)
{
    const u32 ksteps = (n1 * (L + 1) + 31) / 32, groups = (n2 + 3) / 4;
    const size_t gw = (size_t)(L + 1) * c->n * ksteps * IM_WL * 64, words = gw * groups;
    std::vector<u64> key{ n1, n2, g_first, L, c->key_epoch };
    for (u32 k = 1; k < n1; ++k) key.push_back(belt[k]);
    if (t->w_key != key) {
        if (t->w_words < words) {
            cudaFree(t->d_w);
            t->d_w = nullptr;
            t->w_words = 0;
            CU(cudaMalloc(&t->d_w, words * sizeof(u32)));
            t->w_words = words;
        }
        for (u32 g = 0; g < groups; ++g) {
            ImmaPrepParams P{};
            for (u32 k = 1; k < n1; ++k) P.key[k] = c->galois_keys[belt[k]];
            P.diag = t->d;
            P.diag_si = t->stride();
            P.w = t->d_w + g * gw;
            P.n1 = n1;
            P.L = L;
            P.K = c->K;
            P.n = c->n;
            P.g0 = g * 4;
            P.ng = std::min<u32>(4, n2 - g * 4);
            P.ksteps = ksteps;
            dh_imma_prep_kernel<<<c->sms * 8, 256, 0, c->stream>>>(P, c->d_mods);
            c->launches++;
            CU(cudaGetLastError());
        }
        t->w_key = key;
    }
    *w = t->d_w;
    *ksteps_out = ksteps;
    *group_words = gw;
    return HEGPU_OK;
}

// hegpu_matvec_bsgs_range with HEGPU_MATVEC_DH (double-hoisted; oracle: orc_matvec_bsgs_dh)
static int matvec_bsgs_dh(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *in, const hegpu_pt *diags, u32 n1, u32 n2, u32 g_first,
                          bool rescale, const std::vector<u32> &belt, const std::vector<u32> &gelt, bool imma)
{
    const u32 L = in->L, B = in->batch, first_rot = g_first == 0 ? 1u : 0u, nrot = n2 - first_rot;
    const size_t n = c->n, ctw = (size_t)2 * L * n, accw = (size_t)2 * (L + 1) * n;
    if (nrot > (u32)MAXG) INVALID("double-hoisted matvec supports at most 16 rotated giant steps per call");
    const u32 nr1 = std::max<u32>(nrot, 1);
    const bool fused = c->dh_fused && L <= 4 && dh_inner_smem(n1, dh_n2_pad(n2), L) <= (size_t)110 * 1024;
    const size_t nbaby = fused ? 0 : n1;  // the rotated ciphertexts only exist in HBM on the unfused path
    // fused path with more than 4 giant steps (several launches of dh_inner): the first launch keeps the rotated ciphertexts
    // b_k (lazy residues, extended basis) so that the later launches read them back instead of redoing the key products
    const bool keep_bk = fused && c->dh_bk && n2 > 4 && !((imma || c->dh_imma) && L == 3);
    const size_t bkw = keep_bk ? (size_t)(L + 1) * n1 * 2 * n : 0;  // words per ciphertext
    auto need = [&](u32 Bc) {
        return align256((size_t)Bc * bkw) + align256((size_t)Bc * L * n) + align256((size_t)Bc * L * (L + 1) * n) + align256(inv_scratch_words(c, (size_t)Bc * std::max<u32>(L, 2))) +
               align256((size_t)nbaby * Bc * accw) + align256((size_t)n2 * Bc * accw) + align256((size_t)nr1 * Bc * ctw) +
               align256((size_t)nr1 * Bc * 2 * n) + align256(inv_scratch_words(c, (size_t)nr1 * Bc * 2)) + align256((size_t)Bc * accw) +
               align256((size_t)Bc * 2 * n) + 2 * align256((size_t)Bc * ctw) + align256((size_t)Bc * L * n) +
               ks_scratch(c, (size_t)nr1 * Bc, L) + rescale_scratch(c, Bc, 2);
    };
    // two execution slots: alternate batch chunks go to the main and the auxiliary stream (own scratch
    // each), so the small grids of one chunk (a single wave of NTT CTAs) overlap the other chunk's kernels
    const int NS = (!c->profiling && B >= 16u * (u32)c->n_slots) ? c->n_slots : 1;
    u32 Bc = (B + NS - 1) / NS;
    while (Bc > 1 && need(Bc) * NS > c->ws_budget) Bc = (Bc + 1) / 2;
    TRY(arena_reserve(c, need(Bc) * NS));
    const u64 *dmont;
    TRY(pt_montgomery(const_cast<hegpu_pt *>(diags), &dmont));
    imma = (imma || c->dh_imma) && fused && L == 3;  // the MMA form is built for the 3-limb level only
    const u32 *imma_w = nullptr;
    u32 imma_ksteps = 0;
    size_t imma_gw = 0;
    if (imma) {
        TRY(pt_imma_w(c, const_cast<hegpu_pt *>(diags), n1, n2, g_first, L, belt, &imma_w, &imma_ksteps, &imma_gw));
        if (dh_imma_smem(n1, (u32)c->imma_tx) > (size_t)224 * 1024) imma = false;
    }
    SlotGuard guard{ c };
    if (NS > 1) {
        // largest NTT launch of a chunk (the two-level park of N = 32768 parks three quarters per job instead of one half)
        const size_t park_need = (size_t)nr1 * Bc * (L * L + 2 * L) * ((c->logn == 15 && c->park32k == 2) ? 3 * (size_t)(n / 4) : n / 2);
        for (int sl = 0; sl < NS; ++sl) {
            select_slot(c, sl);
            if ((c->logn == 14 || (c->logn == 15 && c->park32k)) && c->use_park) TRY(park_reserve(c, park_need));
        }
        select_slot(c, 0);
        CU(cudaEventRecord(c->ev_fork, c->main_stream));
        for (int sl = 1; sl < NS; ++sl) CU(cudaStreamWaitEvent(c->aux_stream[sl], c->ev_fork, 0));
        guard.forked = NS;
    }
    u32 chunk_no = 0;
    for (u32 b0 = 0; b0 < B; b0 += Bc, ++chunk_no) {
        const u32 Bn = std::min(Bc, B - b0);
        const int slot = (int)(chunk_no % (u32)NS);
        select_slot(c, slot);
        ArenaPlan ap{ c };
        ap.off = (size_t)slot * need(Bc);
        u64 *bk = ap.take((size_t)Bc * bkw);
        u64 *coef = ap.take((size_t)Bc * L * n);
        u64 *ext = ap.take((size_t)Bc * L * (L + 1) * n);
        u64 *scr0 = ap.take(inv_scratch_words(c, (size_t)Bc * std::max<u32>(L, 2)));
        u64 *babyq = ap.take((size_t)nbaby * Bc * accw);
        u64 *u = ap.take((size_t)n2 * Bc * accw);
        u64 *v = ap.take((size_t)nr1 * Bc * ctw);
        u64 *tu = ap.take((size_t)nr1 * Bc * 2 * n);
        u64 *scr1 = ap.take(inv_scratch_words(c, (size_t)nr1 * Bc * 2));
        u64 *accsum = ap.take((size_t)Bc * accw);
        u64 *tsum = ap.take((size_t)Bc * 2 * n);
        u64 *accb = ap.take((size_t)Bc * ctw);
        u64 *c0p = ap.take((size_t)Bc * L * n);
        auto view_of = [&](u64 *p, size_t batch_off) { return CtView{ p + batch_off * ctw, ctw, (size_t)L * n, n }; };
        auto qp_view = [&](u64 *p, size_t batch_off) { return CtView{ p + batch_off * accw, accw, (size_t)(L + 1) * n, n }; };
        const CtView vin = in->view_at(b0);
        // 1. one digit decomposition of c1 for every baby step
        KsPlan pl0;
        pl0.P = KsParams{};
        pl0.P.ngroups = 1;
        pl0.P.B = Bn;
        pl0.P.L = L;
        pl0.P.K = c->K;
        pl0.P.n = c->n;
        pl0.P.target_poly = 1;
        pl0.P.hoisted = 1;
        pl0.P.in[0] = vin;
        pl0.P.coef = coef;
        pl0.P.ext = ext;
        pl0.E = Bn;
        pl0.scr = scr0;
        TRY(ks_decompose(c, pl0));
        if (fused) {
            // 2-4 fused: baby-step key inner products and the inner sums of every giant step in one kernel
            {   // P * c0 once per ciphertext instead of one product per baby step inside dh_inner
                const size_t total = (size_t)Bn * L * n;
                Prof pf(c, PK_ELEMENTWISE, total, total * 16);
                scale_c0_kernel<<<ew_grid(c, total), 256, 0, c->stream>>>(vin, c0p, Bn, L, c->n, c->d_mods);
                c->launches++;
                CU(cudaGetLastError());
            }
            DhInnerParams P{};
            P.in = vin;
            P.c0p = c0p;
            P.ext = ext;
            for (u32 k = 1; k < n1; ++k) {
                P.key[k] = c->galois_keys[belt[k]];
                TRY(get_perm(c, belt[k], &P.perm[k]));
            }
            P.diag = dmont;
            P.diag_si = diags->stride();
            P.u = u;
            P.bk = keep_bk ? bk : nullptr;
            P.n1 = n1;
            P.n2 = n2;
            P.B = Bn;
            P.L = L;
            P.K = c->K;
            P.n = c->n;
            // reads: ciphertexts, lifted digits, keys and diagonals once; writes: the inner sums once
            const u64 words = (u64)Bn * ctw + (u64)Bn * L * (L + 1) * n + (u64)(n1 - 1) * 2 * L * (L + 1) * n +
                              (u64)n1 * n2 * (L + 1) * n + (u64)n2 * Bn * accw;
            Prof pf(c, PK_DH_INNER, (u64)Bn * (L + 1) * n, words * 8);
            if (imma) {
                ImmaParams Q{};
                Q.in = vin;
                Q.ext = ext;
                Q.c0p = c0p;
                for (u32 k = 1; k < n1; ++k) Q.perm[k] = P.perm[k];
                Q.u = u;
                Q.n1 = n1;
                Q.B = Bn;
                Q.L = L;
                Q.K = c->K;
                Q.n = c->n;
                Q.ksteps = imma_ksteps;
                const u32 tx = (u32)c->imma_tx;
                const size_t smem = dh_imma_smem(n1, tx);
                TRY(configure_smem(c, tx == 16 ? (const void *)dh_imma_kernel<3, 16> : (const void *)dh_imma_kernel<3, 8>, smem));
                for (u32 g0 = 0; g0 < n2; g0 += 4) {
                    Q.g0 = g0;
                    Q.ng = std::min<u32>(4, n2 - g0);
                    Q.w = imma_w + (size_t)(g0 / 4) * imma_gw;
                    if (tx == 16)
                        dh_imma_kernel<3, 16><<<(u32)(n / 16) * (L + 1), 512, smem, c->stream>>>(Q, c->d_mods);
                    else
                        dh_imma_kernel<3, 8><<<(u32)(n / 8) * (L + 1), 256, smem, c->stream>>>(Q, c->d_mods);
                    c->launches++;
                    CU(cudaGetLastError());
                }
            } else
#define DH_CASE(LT, N2) case (LT) * 16 + (N2): TRY((launch_dh_inner<LT, N2>(c, P))); break;
            switch (L * 16 + dh_n2_pad(n2)) {
                DH_CASE(1, 1) DH_CASE(1, 2) DH_CASE(1, 4)
                DH_CASE(2, 1) DH_CASE(2, 2) DH_CASE(2, 4)
                DH_CASE(3, 1) DH_CASE(3, 2) DH_CASE(3, 4)
                DH_CASE(4, 1) DH_CASE(4, 2) DH_CASE(4, 4)
            default: LOGIC("unsupported double-hoisted shape");
            }
#undef DH_CASE
        } else {
            // 2. baby step 0 = P * (c0, c1) in the extended basis
            {
                const size_t total = (size_t)Bn * accw;
                Prof pf(c, PK_ELEMENTWISE, total, (size_t)Bn * ctw * 8 + total * 8);
                scale_by_p_kernel<<<ew_grid(c, total), 256, 0, c->stream>>>(vin, babyq, Bn, L, c->n, c->d_mods);
                c->launches++;
                CU(cudaGetLastError());
            }
            // 3. baby steps k >= 1: key inner products of the permuted digits, + P * pi_k(c0); no mod-down
            for (u32 k0 = 1; k0 < n1; k0 += MAXG) {
                KsPlan pk = pl0;
                u32 ng = 0;
                for (u32 k = k0; k < n1 && ng < (u32)MAXG; ++k, ++ng) {
                    const u32 *pm;
                    TRY(get_perm(c, belt[k], &pm));
                    pk.P.in[ng] = vin;
                    pk.P.key[ng] = c->galois_keys[belt[k]];
                    pk.P.perm[ng] = pm;
                }
                pk.P.ngroups = ng;
                pk.P.add_pc0 = 1;
                pk.P.acc = babyq + (size_t)k0 * Bn * accw;
                pk.E = (size_t)ng * Bn;
                TRY(ks_inner(c, pk));
            }
            // 4. inner sums of every giant step in the extended basis
            {
                BsgsParams P{};
                for (u32 k = 0; k < n1; ++k) P.baby[k] = qp_view(babyq, (size_t)k * Bn);
                P.inner = qp_view(u, 0);
                P.diag = dmont;
                P.diag_si = diags->stride();
                P.diag_sl = n;
                P.n1 = n1;
                P.n2 = n2;
                P.B = Bn;
                P.L = L + 1;
                P.n = c->n;
                P.special_limb = L;
                P.K = c->K;
                const size_t total = (size_t)Bn * accw;
                Prof pf(c, PK_BSGS_INNER, total, total * 8 * (n1 + n2) + (u64)n1 * n2 * (L + 1) * n * 8);
                launch_bsgs_inner_n1(c, P, n1);
                c->launches++;
                CU(cudaGetLastError());
            }
        }
        CtView dst = rescale ? view_of(accb, 0) : out->view_at(b0);
        KsPlan pf1;  // the final mod-down
        pf1.P = KsParams{};
        pf1.P.ngroups = 1;
        pf1.P.B = Bn;
        pf1.P.L = L;
        pf1.P.K = c->K;
        pf1.P.n = c->n;
        pf1.P.target_poly = 1;
        pf1.P.out[0] = dst;
        pf1.P.t = tsum;
        pf1.E = Bn;
        pf1.scr = scr1;
        if (nrot == 0) {
            pf1.P.acc = u;  // the single unrotated step
            pf1.P.no_base0 = 1;
            pf1.P.in[0] = dst;
        } else {
            // 5. v1_g = mod_down(u_g[1]) for the rotated giant steps: only the component that is key-switched
            //    leaves the extended basis
            KsPlan pm;
            pm.P = KsParams{};
            pm.P.ngroups = nrot;
            pm.P.B = Bn;
            pm.P.L = L;
            pm.P.K = c->K;
            pm.P.n = c->n;
            pm.P.target_poly = 1;
            pm.P.no_base0 = 1;
            pm.P.only_c1 = 1;
            for (u32 r = 0; r < nrot; ++r) {
                pm.P.in[r] = view_of(v, (size_t)r * Bn);
                pm.P.out[r] = view_of(v, (size_t)r * Bn);
            }
            pm.P.acc = u + (size_t)first_rot * Bn * accw;
            pm.P.t = tu;
            pm.E = (size_t)nrot * Bn;
            pm.scr = scr1;
            TRY(ks_moddown(c, pm));
            // 6. giant rotations: key-switch pi_g(v1_g) without mod-down, summed over g, + pi_g(u_g[0]) (+ u_0)
            KsGroupDesc gs[MAXG];
            for (u32 r = 0; r < nrot; ++r) {
                const u32 g = first_rot + r;
                const u32 *pmr;
                TRY(get_perm(c, gelt[g], &pmr));
                gs[r] = KsGroupDesc{ view_of(v, (size_t)r * Bn), view_of(v, (size_t)r * Bn), c->galois_keys[gelt[g]], pmr };
            }
            KsPlan pl;
            TRY(ks_setup(c, pl, gs, nrot, Bn, L, 1, false, false, ap));
            TRY(ks_decompose(c, pl));
            pl.P.u0 = u + (size_t)first_rot * Bn * accw;
            TRY(ks_inner_sum(c, pl, first_rot ? u : nullptr, accsum));
            pf1.P.acc = accsum;
            pf1.P.no_base0 = 1;
            pf1.P.in[0] = dst;
        }
        if (rescale && c->fuse_final) {
            ArenaPlan ar{ c };
            ar.off = ap.off;
            u64 *t2 = ar.take((size_t)Bn * 2 * n);  // inside the region sized for the unfused rescale
            TRY(ks_moddown_rescale(c, pf1, out->view_at(b0), t2));
        } else {
            TRY(ks_moddown(c, pf1));
            if (rescale) {
                ArenaPlan ar{ c };
                ar.off = ap.off;
                TRY(rescale_views(c, out->view_at(b0), dst, Bn, 2, L, ar));
            }
        }
    }
    return HEGPU_OK;  // SlotGuard joins the auxiliary streams
}

extern "C" int hegpu_matvec_bsgs(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *in, const hegpu_pt *diags, uint32_t n1,
                                 uint32_t n2, int flags)
{
    return hegpu_matvec_bsgs_range(c, out, in, diags, n1, n2, 0, flags);
}

extern "C" int hegpu_matvec_bsgs_range(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *in, const hegpu_pt *diags, uint32_t n1,
                                       uint32_t n2, uint32_t g_first, int flags)
{
    const bool rescale = (flags & HEGPU_MATVEC_RESCALE) != 0, hoist = (flags & HEGPU_MATVEC_HOIST) != 0,
               lazy = (flags & HEGPU_MATVEC_LAZY) != 0, dh = (flags & HEGPU_MATVEC_DH) != 0;
    // giant step g' of this call is global giant step g_first + g'; only global step 0 is unrotated
    const u32 first_rot = g_first == 0 ? 1u : 0u;
    if (!c || !out || !diags) INVALID("null argument");
    TRY(check_ct(in));
    if (in->size != 2) INVALID("encrypted size must be 2");
    if (n1 == 0 || n2 == 0 || n1 > (u32)(dh ? MAXB : MAXG)) INVALID("baby-step count must be in [1,16] ([1,32] double-hoisted)");
    if (dh && !(n1 <= 16 || n1 == 24 || n1 == 32)) INVALID("double-hoisted baby-step count must be <= 16, 24 or 32");
    if (diags->count < n1 * n2 || diags->L != in->L) INVALID("encrypted_ntt and plain_ntt parameter mismatch");
    if (dh != diags->ext) INVALID("HEGPU_MATVEC_DH needs plaintexts uploaded with hegpu_pt_upload_ext (and only DH takes them)");
    if (out == in) INVALID("matvec output must not alias its input");
    const u32 L = in->L, B = in->batch;
    if (rescale && L < 2) INVALID("end of modulus switching chain reached");
    const double ns = in->scale * diags->scale;
    if (!scale_in_bounds(c, ns, L)) INVALID("scale out of bounds");
    TRY(set_device(c));
    TRY(fits(out, B, 2, rescale ? L - 1 : L));
    std::vector<u32> belt(n1, 0), gelt(n2, 0);
    for (u32 k = 1; k < n1; ++k) {
        TRY(hegpu_galois_elt_from_step(c, (int)k, &belt[k]));
        if (!c->galois_keys.count(belt[k])) INVALID("Galois key not present");
    }
    for (u32 g = first_rot; g < n2; ++g) {
        TRY(hegpu_galois_elt_from_step(c, (int)((g_first + g) * n1), &gelt[g]));
        if (!c->galois_keys.count(gelt[g])) INVALID("Galois key not present");
    }
    if (dh) {
        TRY(matvec_bsgs_dh(c, out, in, diags, n1, n2, g_first, rescale, belt, gelt, (flags & HEGPU_MATVEC_IMMA) != 0));
        out->size = 2;
        out->L = rescale ? L - 1 : L;
        out->scale = rescale ? ns / (double)c->q[L - 1] : ns;
        return HEGPU_OK;
    }
    const u32 nrot = n2 - first_rot;  // rotated giant steps
    const size_t n = c->n, ctw = (size_t)2 * L * n;
    // chunk the batch so that the scratch stays within the budget
    auto need = [&](u32 Bc) {
        const u32 gmax = std::min<u32>(MAXG, std::max(n1 - 1, nrot));  // nrot = n2 when g_first > 0
        return ks_scratch(c, (size_t)gmax * Bc, L) + align256((size_t)(n1 - 1) * Bc * ctw) + align256((size_t)n2 * Bc * ctw) +
               align256((size_t)std::max<u32>(nrot, 1) * Bc * ctw) + align256((size_t)Bc * ctw) + rescale_scratch(c, Bc, 2) +
               align256((size_t)Bc * 2 * (L + 1) * n) + align256((size_t)Bc * 2 * n) + align256((size_t)Bc * ctw);
    };
    u32 Bc = B;
    while (Bc > 1 && need(Bc) > c->ws_budget) Bc = (Bc + 1) / 2;
    TRY(arena_reserve(c, need(Bc)));
    for (u32 b0 = 0; b0 < B; b0 += Bc) {
        const u32 Bn = std::min(Bc, B - b0);
        ArenaPlan ap{ c };
        u64 *baby = ap.take((size_t)(n1 - 1) * Bc * ctw);
        u64 *inner = ap.take((size_t)n2 * Bc * ctw);
        u64 *rot = ap.take((size_t)std::max<u32>(nrot, 1) * Bc * ctw);
        u64 *accb = ap.take((size_t)Bc * ctw);
        const size_t ks_off = ap.off;
        auto view_of = [&](u64 *p, size_t batch_off) { return CtView{ p + batch_off * ctw, ctw, (size_t)L * n, n }; };
        const CtView vin = in->view_at(b0);
        // 1. baby steps: all rotations of the input in one fused key-switch launch
        for (u32 k0 = 1; k0 < n1; k0 += MAXG) {
            KsGroupDesc gs[MAXG];
            u32 ng = 0;
            for (u32 k = k0; k < n1 && ng < (u32)MAXG; ++k, ++ng) {
                const u32 *pm;
                TRY(get_perm(c, belt[k], &pm));
                gs[ng] = KsGroupDesc{ vin, view_of(baby, (size_t)(k - 1) * Bn), c->galois_keys[belt[k]], pm };
            }
            ap.off = ks_off;
            TRY(keyswitch(c, gs, ng, Bn, L, 1, false, ap, hoist));
        }
        // 2. inner sums for every giant step
        {
            BsgsParams P{};
            P.baby[0] = vin;
            for (u32 k = 1; k < n1; ++k) P.baby[k] = view_of(baby, (size_t)(k - 1) * Bn);
            P.inner = view_of(inner, 0);
            TRY(pt_montgomery(const_cast<hegpu_pt *>(diags), &P.diag));
            P.diag_si = diags->stride();
            P.diag_sl = n;
            P.n1 = n1;
            P.n2 = n2;
            P.B = Bn;
            P.L = L;
            P.n = c->n;
            P.special_limb = ~0u;
            P.K = c->K;
            const size_t total = (size_t)Bn * ctw;
            // rotated ciphertexts read once, inner sums written once, diagonals read once per step
            Prof pf(c, PK_BSGS_INNER, total, total * 8 * (n1 + n2) + (u64)n1 * n2 * L * n * 8);
            launch_bsgs_inner_n1(c, P, n1);
            c->launches++;
            CU(cudaGetLastError());
        }
        CtView dst = rescale ? view_of(accb, 0) : out->view_at(b0);
        // rotated giant step r (r < nrot) is local step first_rot + r: input inner[first_rot + r], output rot[r]
        if (nrot == 0) {
            TRY(launch_ew<EW_COPY>(c, dst, view_of(inner, 0), view_of(inner, 0), Bn, 2, L));
        } else if (!lazy) {
            // 3. giant steps, 4. accumulate
            for (u32 r0 = 0; r0 < nrot; r0 += MAXG) {
                KsGroupDesc gs[MAXG];
                u32 ng = 0;
                for (u32 r = r0; r < nrot && ng < (u32)MAXG; ++r, ++ng) {
                    const u32 g = first_rot + r;
                    const u32 *pm;
                    TRY(get_perm(c, gelt[g], &pm));
                    gs[ng] = KsGroupDesc{ view_of(inner, (size_t)g * Bn), view_of(rot, (size_t)r * Bn), c->galois_keys[gelt[g]], pm };
                }
                ap.off = ks_off;
                TRY(keyswitch(c, gs, ng, Bn, L, 1, false, ap));
            }
            // sum of inner[0] (if unrotated) and every rot[r]
            SumParams S{ first_rot ? view_of(inner, 0) : view_of(rot, 0), first_rot ? view_of(rot, 0) : view_of(rot, (size_t)Bn), dst,
                         first_rot ? nrot + 1 : nrot, Bn, 2, L, c->n };
            const size_t total = (size_t)Bn * ctw;
            Prof pf(c, PK_ELEMENTWISE, total, total * 8 * (n2 + 1));
            sum_terms_kernel<<<ew_grid(c, total), 256, 0, c->stream>>>(S, c->d_mods);
            c->launches++;
            CU(cudaGetLastError());
        } else {
            // 3'. giant steps with ONE mod-down: the key inner products of all giant steps are summed
            // in the extended basis (mod q_0..q_{L-1}, P) and divided by P once.
            if (nrot > (u32)MAXG) INVALID("lazy giant steps support at most 16 rotated giant steps per call");
            ap.off = ks_off;
            u64 *accsum = ap.take((size_t)Bn * 2 * (L + 1) * n);
            u64 *tsum = ap.take((size_t)Bn * 2 * n);
            u64 *basebuf = ap.take((size_t)Bn * ctw);
            KsGroupDesc gs[MAXG];
            BaseSumParams BS{};
            for (u32 r = 0; r < nrot; ++r) {
                const u32 g = first_rot + r;
                const u32 *pm;
                TRY(get_perm(c, gelt[g], &pm));
                gs[r] = KsGroupDesc{ view_of(inner, (size_t)g * Bn), view_of(rot, (size_t)r * Bn), c->galois_keys[gelt[g]], pm };
                BS.perm[r] = pm;
            }
            KsPlan pl;
            TRY(ks_setup(c, pl, gs, nrot, Bn, L, 1, false, false, ap));
            TRY(ks_decompose(c, pl));
            TRY(ks_inner_sum(c, pl, nullptr, accsum));
            {   // base = [inner_0 if unrotated] + sum_r pi_r(inner_r.c0)
                BS.first = view_of(inner, 0);
                BS.has_first = first_rot;
                BS.rest = view_of(inner, (size_t)first_rot * Bn);
                BS.out = view_of(basebuf, 0);
                BS.groups = nrot;
                BS.B = Bn;
                BS.L = L;
                BS.n = c->n;
                const size_t total = (size_t)Bn * ctw;
                Prof pf(c, PK_ELEMENTWISE, total, total * 8 * 2 + (size_t)Bn * L * n * 8 * nrot);
                base_gather_sum_kernel<<<ew_grid(c, total), 256, 0, c->stream>>>(BS, c->d_mods);
                c->launches++;
                CU(cudaGetLastError());
            }
            KsPlan pm1;
            pm1.P = KsParams{};
            pm1.P.ngroups = 1;
            pm1.P.B = Bn;
            pm1.P.L = L;
            pm1.P.K = c->K;
            pm1.P.n = c->n;
            pm1.P.target_poly = 1;
            pm1.P.has_base1 = 1;
            pm1.P.in[0] = view_of(basebuf, 0);
            pm1.P.out[0] = dst;
            pm1.P.acc = accsum;
            pm1.P.t = tsum;
            pm1.E = Bn;
            pm1.scr = pl.scr;
            TRY(ks_moddown(c, pm1));
        }
        // 5. rescale
        if (rescale) {
            ap.off = ks_off;
            TRY(rescale_views(c, out->view_at(b0), dst, Bn, 2, L, ap));
        }
    }
    out->size = 2;
    out->L = rescale ? L - 1 : L;
    out->scale = rescale ? ns / (double)c->q[L - 1] : ns;
    return HEGPU_OK;
}

// temporaries for the loop-order-exact composites
struct TmpCt {
    hegpu_ct *t = nullptr;  // owned by the context's cache
};
// Temporary batch `slot` of a composite: cached in the context by (slot, batch, size_cap), grown when a deeper level
// is asked for.  Reuse across calls is ordered by the context stream, so the composites never block the host.
static int tmp_ct(hegpu_ctx *c, int slot, hegpu_ct **out, u32 batch, u32 size_cap, u32 L_cap)
{
    hegpu_ct *&t = c->tmp_cache[std::make_tuple(slot, batch, size_cap)];
    if (t && t->L_cap < L_cap) {
        hegpu_ct_destroy(t);
        t = nullptr;
    }
    if (!t) TRY(hegpu_ct_create(c, &t, batch, size_cap, L_cap));
    t->size = 0;
    t->L = 0;
    t->scale = 1.0;
    *out = t;
    return HEGPU_OK;
}

extern "C" int hegpu_bmatmul(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *ths, const hegpu_ct *oth, uint32_t n, uint32_t p,
                             int case_b)
{
    if (!c || !out) INVALID("null argument");
    TRY(check_ct(ths));
    TRY(check_ct(oth));
    if (ths->size != 2 || oth->size != 2) INVALID("encrypted size must be 2");
    if (ths->batch != n || oth->batch != (case_b ? n : p) || out->batch != p) INVALID("batch sizes do not match the matrix shape");
    if (ths->L != oth->L) INVALID("encrypted1 and encrypted2 parameter mismatch");
    if (!c->relin_key) INVALID("not enough relinearization keys");
    if (ths->L < 2) INVALID("end of modulus switching chain reached");
    if (out == ths || out == oth) INVALID("output must not alias an input");
    TRY(set_device(c));
    const u32 L = ths->L;
    TmpCt acc, rot, prod, one;
    TRY(tmp_ct(c, 0, &acc.t, p, 3, L));
    if (!case_b) {
        // case A: res_i = sum_j rot(other_i, j) * this_j; batch over i, `this_j` broadcast
        TRY(tmp_ct(c, 1, &rot.t, p, 2, L));
        TRY(tmp_ct(c, 2, &prod.t, p, 3, L));
        TRY(tmp_ct(c, 3, &one.t, 1, 2, L));
        for (u32 j = 0; j < n; ++j) {
            TRY(hegpu_rotate_vector(c, rot.t, oth, (int)j));
            TRY(hegpu_ct_copy_one(c, one.t, 0, ths, j));
            one.t->scale = ths->scale;
            if (j == 0) {
                TRY(hegpu_multiply(c, acc.t, rot.t, one.t));
            } else {
                TRY(hegpu_multiply(c, prod.t, rot.t, one.t));
                TRY(hegpu_add(c, acc.t, acc.t, prod.t));
            }
        }
    } else {
        // case B: res_i = sum_j rot(other_j, i) * this_j; batch over j, summed over the batch
        TRY(tmp_ct(c, 1, &rot.t, n, 2, L));
        TRY(tmp_ct(c, 2, &prod.t, n, 3, L));
        for (u32 i = 0; i < p; ++i) {
            TRY(hegpu_rotate_vector(c, rot.t, oth, (int)i));
            TRY(hegpu_multiply(c, prod.t, rot.t, ths));
            CtView first = prod.t->view_at(0), rest = prod.t->view_at(n > 1 ? 1 : 0);
            SumParams S{ first, rest, acc.t->view_at(i), n, 1, 3, L, c->n };
            const size_t total = (size_t)3 * L * c->n;
            Prof pf(c, PK_ELEMENTWISE, total, total * 8 * (n + 1));
            sum_terms_kernel<<<ew_grid(c, total), 256, 0, c->stream>>>(S, c->d_mods);
            c->launches++;
            CU(cudaGetLastError());
            acc.t->size = 3;
            acc.t->L = L;
            acc.t->scale = prod.t->scale;
        }
    }
    TRY(hegpu_relinearize(c, acc.t, acc.t));
    TRY(hegpu_rescale_to_next(c, out, acc.t));
    return HEGPU_OK;
}

// out(i,j) index in a column-major rows x cols matrix with the reference's transposed flag
// (he_linalg.cpp:376-379: ij_to_idx = transposed ? i*cols' ... ) -- see host/he_linalg for
// the mirrored class; here: idx = transposed ? (i * cols + j) : (i + j * rows).
static inline u32 mat_idx(u32 i, u32 j, u32 rows, u32 cols, int transposed) { return transposed ? i * cols + j : i + j * rows; }

__global__ void __launch_bounds__(256) matmul_tensor_kernel(CtView out, CtView a, CtView b, u32 rows, u32 inner, u32 cols,
                                                            int at, int bt, u32 L, u32 n, const ModConst *__restrict__ mods)
{
    const size_t per_o = (size_t)L * n;
    const size_t total = (size_t)rows * cols * per_o;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const u32 x = (u32)(idx % n);
        size_t r = idx / n;
        const u32 l = (u32)(r % L);
        const u32 o = (u32)(r / L);  // column-major output index
        const u32 i = o % rows, j = o / rows;
        const ModConst m = mods[l];
        u64 h0 = 0, l0 = 0, h1 = 0, l1 = 0, h2 = 0, l2 = 0;
        for (u32 k = 0; k < inner; ++k) {
            const u32 ia = at ? i * inner + k : i + k * rows;
            const u32 ib = bt ? k * cols + j : k + j * inner;
            const u64 *pa = a.p + ia * a.sb + l * a.sl + x, *pb = b.p + ib * b.sb + l * b.sl + x;
            const u64 a0 = pa[0], a1 = pa[a.sp], b0 = pb[0], b1 = pb[b.sp];
            mac128(h0, l0, a0, b0);
            mac128(h1, l1, a0, b1);
            mac128(h1, l1, a1, b0);
            mac128(h2, l2, a1, b1);
            if ((k & 31) == 31) {  // keep the lazy 128-bit sums far from overflow
                l0 = barrett128(h0, l0, m); h0 = 0;
                l1 = barrett128(h1, l1, m); h1 = 0;
                l2 = barrett128(h2, l2, m); h2 = 0;
            }
        }
        u64 *po = out.p + o * out.sb + l * out.sl + x;
        po[0] = barrett128(h0, l0, m);
        po[out.sp] = barrett128(h1, l1, m);
        po[2 * out.sp] = barrett128(h2, l2, m);
    }
}

extern "C" int hegpu_matmul_elemwise(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *a, const hegpu_ct *b, uint32_t rows,
                                     uint32_t inner, uint32_t cols, int at, int bt)
{
    if (!c || !out) INVALID("null argument");
    TRY(check_ct(a));
    TRY(check_ct(b));
    if (a->size != 2 || b->size != 2) INVALID("encrypted size must be 2");
    if (a->batch != rows * inner || b->batch != inner * cols || out->batch != rows * cols) INVALID("batch sizes do not match the matrix shape");
    if (a->L != b->L) INVALID("encrypted1 and encrypted2 parameter mismatch");
    if (!c->relin_key) INVALID("not enough relinearization keys");
    if (a->L < 2) INVALID("end of modulus switching chain reached");
    if (out == a || out == b) INVALID("output must not alias an input");
    const double ns = a->scale * b->scale;
    if (!scale_in_bounds(c, ns, a->L)) INVALID("scale out of bounds");
    TRY(set_device(c));
    TmpCt acc;
    TRY(tmp_ct(c, 0, &acc.t, rows * cols, 3, a->L));
    const size_t total = (size_t)rows * cols * a->L * c->n;
    Prof pf(c, PK_TENSOR, total, total * 8 * (4 * inner + 3));
    matmul_tensor_kernel<<<ew_grid(c, total), 256, 0, c->stream>>>(acc.t->view(), a->view(), b->view(), rows, inner, cols, at, bt,
                                                                    a->L, c->n, c->d_mods);
    c->launches++;
    CU(cudaGetLastError());
    acc.t->size = 3;
    acc.t->L = a->L;
    acc.t->scale = ns;
    (void)mat_idx;
    TRY(hegpu_relinearize(c, acc.t, acc.t));
    TRY(hegpu_rescale_to_next(c, out, acc.t));
    return HEGPU_OK;
}

extern "C" int hegpu_bfft_stage(hegpu_ctx *c, hegpu_ct *y, const hegpu_pt *pts, int steps, int with_d2)
{
    if (!c || !pts) INVALID("null argument");
    TRY(check_ct(y));
    if (pts->count < (with_d2 ? 3u : 2u)) INVALID("stage needs the D0, D1 (and D2) plaintexts");
    TmpCt y0, y1, y2;
    TRY(tmp_ct(c, 4, &y0.t, y->batch, 2, y->L));
    TRY(tmp_ct(c, 5, &y1.t, y->batch, 2, y->L));
    TRY(hegpu_multiply_plain(c, y0.t, y, pts, 0));
    TRY(hegpu_rescale_to_next(c, y0.t, y0.t));
    TRY(hegpu_rotate_vector(c, y1.t, y, steps));
    TRY(hegpu_multiply_plain(c, y1.t, y1.t, pts, 1));
    TRY(hegpu_rescale_to_next(c, y1.t, y1.t));
    if (with_d2) {
        TRY(tmp_ct(c, 6, &y2.t, y->batch, 2, y->L));
        TRY(hegpu_rotate_vector(c, y2.t, y, -steps));
        TRY(hegpu_multiply_plain(c, y2.t, y2.t, pts, 2));
        TRY(hegpu_rescale_to_next(c, y2.t, y2.t));
    }
    TRY(hegpu_add(c, y, y0.t, y1.t));
    if (with_d2) TRY(hegpu_add(c, y, y, y2.t));
    return HEGPU_OK;
}

extern "C" int hegpu_fft_butterflies(hegpu_ctx *c, hegpu_ct *out, const hegpu_ct *even, const hegpu_ct *odd,
                                     const hegpu_pt *w, const hegpu_pt *one)
{
    if (!c || !out || !w || !one) INVALID("null argument");
    TRY(check_ct(even));
    TRY(check_ct(odd));
    const u32 half = even->batch;
    if (odd->batch != half || out->batch != 2 * half || w->count != half) INVALID("batch sizes do not match");
    if (out == even || out == odd) INVALID("output must not alias an input");
    TmpCt t, e;
    TRY(tmp_ct(c, 7, &t.t, half, 2, odd->L));
    TRY(tmp_ct(c, 8, &e.t, half, 2, even->L));
    TRY(hegpu_multiply_plain(c, t.t, odd, w, -1));
    TRY(hegpu_rescale_to_next(c, t.t, t.t));
    TRY(hegpu_multiply_plain(c, e.t, even, one, 0));
    TRY(hegpu_rescale_to_next(c, e.t, e.t));
    if (!are_close(e.t->scale, t.t->scale)) INVALID("scale mismatch");
    const u32 L = e.t->L;
    if (L > out->L_cap) INVALID("destination capacity too small");
    CtView ve = e.t->view(), vt = t.t->view();
    TRY(launch_ew<EW_ADD>(c, out->view_at(0), ve, vt, half, 2, L));
    TRY(launch_ew<EW_SUB>(c, out->view_at(half), ve, vt, half, 2, L));
    out->size = 2;
    out->L = L;
    out->scale = e.t->scale;
    return HEGPU_OK;
}
