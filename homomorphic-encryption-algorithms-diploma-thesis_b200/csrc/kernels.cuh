// kernels.cuh -- the CUDA kernels of the CKKS evaluator (sm_100a).
//
//  K1/K2  ntt_fwd_kernel / ntt_inv_kernel (+ ntt_inv_final_kernel for N = 32768)
//  K3-K5  element-wise: add / sub / negate / multiply_plain / tensor product / square
//  K6     Galois permutation: fused as a gather into the key-switch INTT load, the key
//         inner product (digit i == j) and the mod-down epilogue -- never a separate pass
//  K7     key-switch: ks INTT -> lift+NTT -> key inner product -> INTT(+half) -> mod-down NTT
//  K8     rescale: INTT(+half) -> mod-down NTT (same two kernels as the end of K7)
//  K10    bsgs_inner_kernel: sum_k baby_k (.) diag_{g,k} for all giant steps at once
//  K11    fixup_kernel: reduce NCCL uint64 sums back to [0,q)
#pragma once
#include "ntt.cuh"

namespace hegpu {

constexpr int MAXG = 16;  // rotation groups fused into one key-switch launch
constexpr int MAXB = 32;  // baby steps of one BSGS matvec

// strided view of a ciphertext batch in HBM: word (b, p, l, x) at p + b*sb + p*sp + l*sl + x
struct CtView {
    u64 *p;
    size_t sb, sp, sl;
};

// dropped-modulus constants: d^-1 mod q_i and floor(d/2) mod q_i
struct MdConst {
    u64 inv, inv_sh, halfmod, pad;
};

// ---------------------------------------------------------------------------------------
// NTT job resolvers.  A resolver maps a job id (blockIdx.x >> SPLIT) to the modulus and to
// the load / store functions of that limb polynomial.
// ---------------------------------------------------------------------------------------

// plain transform of `count` limb polynomials [count][N] (measurement API, host tooling)
struct PlainJob {
    static constexpr bool PIPE = true;  // software-pipelined first-pass loads (no epilogue operands -> no spills)
    const u64 *src;
    u64 *dst;
    u32 first_mod, n_mods, n;
    __device__ __forceinline__ u32 mod(u32 j) const { return first_mod + j % n_mods; }
    __device__ __forceinline__ u64 load_raw(u32 j, u32 i) const { return src[(size_t)j * n + i]; }
    __device__ __forceinline__ u64 load_fix(u32, u64 v, const ModConst &) const { return v; }
    __device__ __forceinline__ void store(u32 j, u32 i, u64 v, const ModConst &) const { dst[(size_t)j * n + i] = v; }
    struct Ops {};
    __device__ __forceinline__ Ops fetch(u32, u32, const ModConst &) const { return Ops{}; }
    __device__ __forceinline__ void store(u32 j, u32 i, u64 v, const ModConst &m, const Ops &) const { store(j, i, v, m); }
};

// key-switch launch parameters (one launch handles ngroups rotations/relinearisations of
// B ciphertexts each: element e = g*B + b)
struct KsParams {
    u32 ngroups, B, L, K, n;
    u32 target_poly;  // 1: Galois (switch pi(c1)); 2: relinearise (switch c2)
    u32 has_base1;    // add in[1] to component 1 (relinearise)
    u32 hoisted;      // all groups rotate the same input: INTT + lift run once per ciphertext (no
                      // gather) and the Galois permutation is applied to the lifted digits instead
    u32 add_pc0;      // double-hoisted baby steps: acc[0][i] += P * pi(c0)[i] (i < L), the rotated
                      // ciphertext stays in the extended basis scaled by P
    u32 no_base0;     // mod-down without a base ciphertext for component 0 (double-hoisted inner sums)
    CtView in[MAXG];
    CtView out[MAXG];
    const u64 *key[MAXG];   // [Lmax][2][K][N]
    const u32 *perm[MAXG];  // Galois gather table or null
    u64 *coef;              // [E][L][N]
    u64 *ext;               // [E][L][L+1][N]
    u64 *acc;               // [E][2][L+1][N]
    u64 *t;                 // [E][2][N]
};

// K7 step 1: c_j = INTT_{q_j}( pi(target)[j] )
struct KsInttJob {
    static constexpr bool PIPE = false;
    KsParams P;
    __device__ __forceinline__ u32 mod(u32 j) const { return j % P.L; }
    __device__ __forceinline__ u64 load_raw(u32 j, u32 i) const
    {
        const u32 e = j / P.L, l = j % P.L, g = e / P.B, b = e % P.B;
        const CtView &v = P.in[g];
        const u32 *pm = P.hoisted ? nullptr : P.perm[g];
        const u32 src = pm ? __ldg(pm + i) : i;
        return v.p[b * v.sb + P.target_poly * v.sp + l * v.sl + src];
    }
    __device__ __forceinline__ u64 load_fix(u32, u64 v, const ModConst &) const { return v; }
    __device__ __forceinline__ void store(u32 j, u32 i, u64 x, const ModConst &) const { P.coef[(size_t)j * P.n + i] = x; }
};

// K7 step 2: ext[e][j][i] = NTT_{m_i}( c_j mod m_i ), i != j, i in [0, L]
struct KsLiftJob {
    static constexpr bool PIPE = false;
    KsParams P;
    const ModConst *mods;
    __device__ __forceinline__ void split(u32 j, u32 &e, u32 &dj, u32 &di) const
    {
        const u32 LL = P.L * P.L;
        e = j / LL;
        const u32 r = j % LL;
        dj = r / P.L;
        const u32 ii = r % P.L;
        di = ii < dj ? ii : ii + 1;
    }
    __device__ __forceinline__ u32 mod(u32 j) const
    {
        u32 e, dj, di;
        split(j, e, dj, di);
        return di == P.L ? P.K - 1 : di;
    }
    __device__ __forceinline__ u64 load_raw(u32 j, u32 i) const
    {
        u32 e, dj, di;
        split(j, e, dj, di);
        return P.coef[((size_t)e * P.L + dj) * P.n + i];
    }
    __device__ __forceinline__ u64 load_fix(u32 j, u64 v, const ModConst &m) const
    {
        u32 e, dj, di;
        split(j, e, dj, di);
        return (mods[dj].q > m.q) ? barrett64(v, m) : v;
    }
    __device__ __forceinline__ void store(u32 j, u32 i, u64 x, const ModConst &) const
    {
        u32 e, dj, di;
        split(j, e, dj, di);
        P.ext[(((size_t)e * P.L + dj) * (P.L + 1) + di) * P.n + i] = x;
    }
    struct Ops {};
    __device__ __forceinline__ Ops fetch(u32, u32, const ModConst &) const { return Ops{}; }
    __device__ __forceinline__ void store(u32 j, u32 i, u64 v, const ModConst &m, const Ops &) const { store(j, i, v, m); }
};

// INTT of a dropped limb with the rounding offset added: t = (INTT_d(src) + floor(d/2)) mod d.
// Used by K7 step 4 (d = special prime) and by rescale (d = q_{L-1}).
struct HalfInttJob {
    static constexpr bool PIPE = false;
    const u64 *src;  // job j at src + (j / inner) * s_outer + (j % inner) * s_inner
    u64 *dst;        // [jobs][N]
    size_t s_outer, s_inner;
    u32 inner, drop_mod, n;
    __device__ __forceinline__ u32 mod(u32) const { return drop_mod; }
    __device__ __forceinline__ u64 load_raw(u32 j, u32 i) const { return src[(j / inner) * s_outer + (j % inner) * s_inner + i]; }
    __device__ __forceinline__ u64 load_fix(u32, u64 v, const ModConst &) const { return v; }
    __device__ __forceinline__ void store(u32 j, u32 i, u64 x, const ModConst &m) const
    {
        dst[(size_t)j * n + i] = addmod(x, m.q >> 1, m.q);
    }
};

// K7 step 5: out[c][i] = base_c[i] + (acc[c][i] - NTT_{q_i}((t_c mod q_i) - half)) * P^-1
struct KsModDownJob {
    static constexpr bool PIPE = false;
    KsParams P;
    const MdConst *md;  // [K] constants of the dropped modulus (special prime) per target limb
    const ModConst *mods;
    __device__ __forceinline__ void split(u32 j, u32 &e, u32 &c, u32 &l) const
    {
        l = j % P.L;
        c = (j / P.L) & 1u;
        e = j / (2 * P.L);
    }
    __device__ __forceinline__ u32 mod(u32 j) const { return j % P.L; }
    __device__ __forceinline__ u64 load_raw(u32 j, u32 i) const { return P.t[(size_t)(j / P.L) * P.n + i]; }
    __device__ __forceinline__ u64 load_fix(u32 j, u64 v, const ModConst &m) const
    {
        if (mods[P.K - 1].q > m.q) v = barrett64(v, m);
        return submod(v, md[j % P.L].halfmod, m.q);
    }
    struct Ops {
        u64 acc, base;
    };
    __device__ __forceinline__ Ops fetch(u32 j, u32 i, const ModConst &) const
    {
        u32 e, c, l;
        split(j, e, c, l);
        const u32 g = e / P.B, b = e % P.B;
        Ops o;
        o.acc = P.acc[(((size_t)e * 2 + c) * (P.L + 1) + l) * P.n + i];
        const CtView &vi = P.in[g];
        o.base = 0;
        if (c == 0) {
            if (!P.no_base0) {
                const u32 *pm = P.perm[g];
                o.base = vi.p[b * vi.sb + l * vi.sl + (pm ? __ldg(pm + i) : i)];
            }
        } else if (P.has_base1) {
            o.base = vi.p[b * vi.sb + vi.sp + l * vi.sl + i];
        }
        return o;
    }
    __device__ __forceinline__ void store(u32 j, u32 i, u64 x, const ModConst &m, const Ops &o) const
    {
        u32 e, c, l;
        split(j, e, c, l);
        const u32 g = e / P.B, b = e % P.B;
        const u64 r = mul_shoup(submod(o.acc, x, m.q), md[l].inv, md[l].inv_sh, m.q);
        const CtView &vo = P.out[g];
        vo.p[b * vo.sb + c * vo.sp + l * vo.sl + i] = addmod(o.base, r, m.q);
    }
};

// rescale step 2: out[b][p][i] = (a[b][p][i] - NTT_{q_i}((t mod q_i) - half)) * q_last^-1
struct RescaleJob {
    static constexpr bool PIPE = false;
    CtView a, out;
    const u64 *t;       // [B*size][N]
    const MdConst *md;  // constants of dropped modulus q_{L-1} per target limb
    const ModConst *mods;
    u32 size, Lm1, drop_mod, n;  // Lm1 = L-1 target limbs
    __device__ __forceinline__ u32 mod(u32 j) const { return j % Lm1; }
    __device__ __forceinline__ u64 load_raw(u32 j, u32 i) const { return t[(size_t)(j / Lm1) * n + i]; }
    __device__ __forceinline__ u64 load_fix(u32 j, u64 v, const ModConst &m) const
    {
        if (mods[drop_mod].q > m.q) v = barrett64(v, m);
        return submod(v, md[j % Lm1].halfmod, m.q);
    }
    struct Ops {
        u64 av;
    };
    __device__ __forceinline__ Ops fetch(u32 j, u32 i, const ModConst &) const
    {
        const u32 l = j % Lm1, bp = j / Lm1, p = bp % size, b = bp / size;
        return Ops{ a.p[b * a.sb + p * a.sp + l * a.sl + i] };
    }
    __device__ __forceinline__ void store(u32 j, u32 i, u64 x, const ModConst &m, const Ops &o) const
    {
        const u32 l = j % Lm1, bp = j / Lm1, p = bp % size, b = bp / size;
        out.p[b * out.sb + p * out.sp + l * out.sl + i] = mul_shoup(submod(o.av, x, m.q), md[l].inv, md[l].inv_sh, m.q);
    }
};

// ---------------------------------------------------------------------------------------
// Loaders (protocol in ntt.cuh): raw() = memory only, fix() = arithmetic only.
// ---------------------------------------------------------------------------------------
template <class Job>
struct PlainLoader {  // coefficient boff + i of job jid
    typedef u64 Raw;
    static constexpr bool PIPE = Job::PIPE;
    const Job &job;
    const ModConst &m;
    u32 jid, boff;
    __device__ __forceinline__ Raw raw(u32 i) const { return job.load_raw(jid, boff + i); }
    __device__ __forceinline__ u64 fix(Raw r, u32) const { return job.load_fix(jid, r, m); }
};
struct Pair64 {
    u64 x, y;
};
// first (stride-N/2) forward stage folded into the load: returns the half selected by `h`
// (SPLIT kernels) and, when park != null, stores the other half there (park kernels, h = 0)
template <class Job, int LOGL>
struct FoldLoader {
    typedef Pair64 Raw;
    static constexpr bool PIPE = Job::PIPE;
    const Job &job;
    const ModConst &m;
    u32 jid, h;
    ulonglong2 W;  // twiddle of the first stage
    u64 *park;
    __device__ __forceinline__ Raw raw(u32 i) const { return Pair64{ job.load_raw(jid, i), job.load_raw(jid, i + (1u << LOGL)) }; }
    __device__ __forceinline__ u64 fix(Raw r, u32 i) const
    {
        const u64 X = job.load_fix(jid, r.x, m), Y = job.load_fix(jid, r.y, m);
        const u64 Tm = mul_shoup_lazy(Y, W.x, W.y, m.q);
        const u64 top = X + Tm, bot = X + (m.q << 1) - Tm;
        if (park) park[i] = h ? top : bot;
        return h ? bot : top;
    }
};
template <bool PIPE_>
struct ParkLoader {  // second half of a park kernel: what FoldLoader parked (same thread wrote it)
    typedef u64 Raw;
    static constexpr bool PIPE = PIPE_;
    const u64 *park;
    __device__ __forceinline__ Raw raw(u32 i) const { return park[i]; }
    __device__ __forceinline__ u64 fix(Raw r, u32) const { return r; }
};

// ---------------------------------------------------------------------------------------
// NTT kernels.  grid = jobs << SPLIT, dynamic smem = 8 << LOGL.
// ---------------------------------------------------------------------------------------
template <int LOGL, int SPLIT, int LOGE, class Job>
__global__ void __launch_bounds__(NttShape<LOGL, LOGE>::THREADS, NttShape<LOGL, LOGE>::MINB) ntt_fwd_kernel(const Job job, const NttTables T)
{
    extern __shared__ __align__(16) u64 sm[];
    const u32 jid = blockIdx.x >> SPLIT;
    const u32 h = blockIdx.x & ((1u << SPLIT) - 1u);
    const u32 mi = job.mod(jid);
    const ModConst m = T.mods[mi];
    const ulonglong2 *tw = T.fwd + (size_t)mi * T.n;
    const u32 boff = h << LOGL;
    auto fetch = [&](u32 i) { return job.fetch(jid, boff + i, m); };
    auto store = [&](u32 i, u64 v, const typename Job::Ops &o) { job.store(jid, boff + i, v, m, o); };
    const ulonglong2 nowl = make_ulonglong2(0, 0);
    if constexpr (SPLIT == 0) {
        PlainLoader<Job> load{ job, m, jid, 0 };
        if (m.big & 4u)
            ntt_fwd_cta<LOGL, LOGE>(load, fetch, store, T.fwd_d + (size_t)mi * T.n, T.n, ArF64(T.modsd[mi]), sm);
        else if (m.big & 1u)
            ntt_fwd_cta<LOGL, LOGE>(load, fetch, store, tw, T.n, ArI64<true>(m, nowl), sm);
        else
            ntt_fwd_cta<LOGL, LOGE>(load, fetch, store, tw, T.n, ArI64<false>(m, nowl), sm);
    } else {
        // stage 1 (stride N/2) redone from global memory by both halves
        FoldLoader<Job, LOGL> load{ job, m, jid, h, __ldg(tw + 1), nullptr };
        if (m.big & 1u)
            ntt_fwd_cta<LOGL, LOGE>(load, fetch, store, tw, T.n + boff, ArI64<true>(m, nowl), sm);
        else
            ntt_fwd_cta<LOGL, LOGE>(load, fetch, store, tw, T.n + boff, ArI64<false>(m, nowl), sm);
    }
}

// SPLIT = 1: the CTA transforms its half and leaves lazy values in `scratch`
// ([jobs][N]); ntt_inv_final_kernel applies the last stage and the job's store.
template <int LOGL, int SPLIT, int LOGE, class Job>
__global__ void __launch_bounds__(NttShape<LOGL, LOGE>::THREADS, NttShape<LOGL, LOGE>::MINB)
    ntt_inv_kernel(const Job job, const NttTables T, u64 *__restrict__ scratch)
{
    extern __shared__ __align__(16) u64 sm[];
    const u32 jid = blockIdx.x >> SPLIT;
    const u32 h = blockIdx.x & ((1u << SPLIT) - 1u);
    const u32 mi = job.mod(jid);
    const ModConst m = T.mods[mi];
    const ulonglong2 *tw = T.inv + (size_t)mi * T.n;
    const ulonglong2 wl = T.inv_last[mi];
    const u32 boff = h << LOGL;
    PlainLoader<Job> load{ job, m, jid, boff };
    if constexpr (SPLIT == 0) {
        auto store = [&](u32 i, u64 v) { job.store(jid, i, v, m); };
        if (m.big & 4u)
            ntt_inv_cta<LOGL, LOGE, LOGL - 1>(load, store, T.inv_d + (size_t)mi * T.n, T.n, ArF64(T.modsd[mi]), sm);
        else if (m.big & 2u)
            ntt_inv_cta<LOGL, LOGE, LOGL - 1>(load, store, tw, T.n, ArI64<true>(m, wl), sm);
        else
            ntt_inv_cta<LOGL, LOGE, LOGL - 1>(load, store, tw, T.n, ArI64<false>(m, wl), sm);
    } else {
        u64 *dst = scratch + (size_t)jid * T.n + boff;
        auto store = [&](u32 i, u64 v) { dst[i] = v; };
        if (m.big & 2u)
            ntt_inv_cta<LOGL, LOGE, -1>(load, store, tw, T.n + boff, ArI64<true>(m, wl), sm);
        else
            ntt_inv_cta<LOGL, LOGE, -1>(load, store, tw, T.n + boff, ArI64<false>(m, wl), sm);
    }
}


// ---------------------------------------------------------------------------------------
// "park" kernels: N = 2^(LOGL+1) transformed by ONE CTA in 2^LOGL words of shared memory.
// Forward: the stride-N/2 stage is computed while loading half 0; its bottom outputs are parked
// in `park` ([jobs][N/2], written and re-read by the same thread, L2-resident) and transformed
// second.  Inverse: half 0 is transformed and parked, half 1 is transformed and its last-pass
// registers are combined with the parked half in the final stride-N/2 stage.
// ---------------------------------------------------------------------------------------
template <int LOGL, int LOGE, class Job>
__global__ void __launch_bounds__(NttShape<LOGL, LOGE>::THREADS, NttShape<LOGL, LOGE>::MINB)
    ntt_fwd_park_kernel(const Job job, const NttTables T, u64 *__restrict__ park)
{
    extern __shared__ __align__(16) u64 sm[];
    const u32 jid = blockIdx.x;
    const u32 mi = job.mod(jid);
    const ModConst m = T.mods[mi];
    const ulonglong2 *tw = T.fwd + (size_t)mi * T.n;
    constexpr u32 half = 1u << LOGL;
    u64 *pk = park + (size_t)jid * half;
    FoldLoader<Job, LOGL> load0{ job, m, jid, 0, __ldg(tw + 1), pk };
    ParkLoader<Job::PIPE> load1{ pk };
    auto fetch0 = [&](u32 i) { return job.fetch(jid, i, m); };
    auto fetch1 = [&](u32 i) { return job.fetch(jid, half + i, m); };
    auto store0 = [&](u32 i, u64 v, const typename Job::Ops &o) { job.store(jid, i, v, m, o); };
    auto store1 = [&](u32 i, u64 v, const typename Job::Ops &o) { job.store(jid, half + i, v, m, o); };
    const ulonglong2 nowl = make_ulonglong2(0, 0);
    if (m.big & 4u) {
        const ArF64 ar(T.modsd[mi]);
        const double *twd = T.fwd_d + (size_t)mi * T.n;
        ntt_fwd_cta<LOGL, LOGE>(load0, fetch0, store0, twd, T.n, ar, sm);
        __syncthreads();
        ntt_fwd_cta<LOGL, LOGE>(load1, fetch1, store1, twd, T.n + half, ar, sm);
    } else if (m.big & 1u) {
        const ArI64<true> ar(m, nowl);
        ntt_fwd_cta<LOGL, LOGE>(load0, fetch0, store0, tw, T.n, ar, sm);
        __syncthreads();
        ntt_fwd_cta<LOGL, LOGE>(load1, fetch1, store1, tw, T.n + half, ar, sm);
    } else {
        const ArI64<false> ar(m, nowl);
        ntt_fwd_cta<LOGL, LOGE>(load0, fetch0, store0, tw, T.n, ar, sm);
        __syncthreads();
        ntt_fwd_cta<LOGL, LOGE>(load1, fetch1, store1, tw, T.n + half, ar, sm);
    }
}

template <int LOGL, int LOGE, class A, class TWP, class Job>
__device__ __forceinline__ void inv_park_body(const Job &job, u32 jid, const ModConst &m, const A &ar, const TWP tw, u32 n,
                                              typename A::V *pk, u64 *sm)
{
    constexpr u32 half = 1u << LOGL;
    PlainLoader<Job> load0{ job, m, jid, 0 }, load1{ job, m, jid, half };
    auto store0 = [&](u32 i, typename A::V v) { pk[i] = v; };
    auto store1 = [&](u32 i, typename A::V y) {
        typename A::V x = pk[i];
        ar.template inv_bfly_last<LOGL>(x, y);
        job.store(jid, i, ar.inv_final(x), m);
        job.store(jid, half + i, ar.inv_final(y), m);
    };
    ntt_inv_cta<LOGL, LOGE, -1>(load0, store0, tw, n, ar, sm);
    __syncthreads();
    ntt_inv_cta<LOGL, LOGE, -1>(load1, store1, tw, n + half, ar, sm);
}

template <int LOGL, int LOGE, class Job>
__global__ void __launch_bounds__(NttShape<LOGL, LOGE>::THREADS, NttShape<LOGL, LOGE>::MINB)
    ntt_inv_park_kernel(const Job job, const NttTables T, u64 *__restrict__ park)
{
    extern __shared__ __align__(16) u64 sm[];
    const u32 jid = blockIdx.x;
    const u32 mi = job.mod(jid);
    const ModConst m = T.mods[mi];
    const ulonglong2 *tw = T.inv + (size_t)mi * T.n;
    u64 *pk = park + ((size_t)jid << LOGL);
    if (m.big & 4u)
        inv_park_body<LOGL, LOGE>(job, jid, m, ArF64(T.modsd[mi]), T.inv_d + (size_t)mi * T.n, T.n, reinterpret_cast<double *>(pk), sm);
    else if (m.big & 2u)
        inv_park_body<LOGL, LOGE>(job, jid, m, ArI64<true>(m, T.inv_last[mi]), tw, T.n, pk, sm);
    else
        inv_park_body<LOGL, LOGE>(job, jid, m, ArI64<false>(m, T.inv_last[mi]), tw, T.n, pk, sm);
}

// last (stride N/2) INTT stage for N = 2^(LOGL+1), element-wise over scratch
template <int LOGL, class Job>
__global__ void __launch_bounds__(256) ntt_inv_final_kernel(const Job job, const NttTables T, const u64 *__restrict__ scratch,
                                                            u32 jobs)
{
    const u32 halfn = 1u << LOGL;
    const size_t total = (size_t)jobs * halfn;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const u32 jid = (u32)(idx >> LOGL), i = (u32)(idx & (halfn - 1));
        const u32 mi = job.mod(jid);
        const ModConst m = T.mods[mi];
        const ulonglong2 wl = T.inv_last[mi];
        const u64 X = scratch[(size_t)jid * T.n + i], Y = scratch[(size_t)jid * T.n + halfn + i];
        u64 Sm, D;
        if (m.big & 2u) {
            Sm = csub(X + Y, m.q << 1);
            D = X + (m.q << 1) - Y;
        } else {
            Sm = X + Y;
            D = X + (m.q << (LOGL + 1)) - Y;
        }
        job.store(jid, i, mul_shoup(Sm, m.ninv, m.ninv_sh, m.q), m);
        job.store(jid, i + halfn, mul_shoup(D, wl.x, wl.y, m.q), m);
    }
}

// ---------------------------------------------------------------------------------------
// K7 step 3: key inner product.  acc[e][c][i] = sum_j ext[e][j][i] (.) key[j][c][ki] mod m_i;
// digit i == j reads the (permuted) target directly.  128-bit lazy accumulation, one
// reduction (SEAL: Barrett; here the device copy of every key is kept in Montgomery form
// k*2^64 mod m, so one Montgomery reduction returns the same canonical residue at a third of
// the multiplies).  A thread owns one (group, ext limb, coefficient), keeps the
// 2L key words of that coefficient in registers and sweeps a chunk of the batch, so the keys
// are read once per chunk instead of once per ciphertext.  grid = (x blocks, ngroups*(L+1),
// batch chunks).  LT = L when L <= 4 (keys in registers), 0 = generic.
// ---------------------------------------------------------------------------------------
constexpr int KS_INNER_BCHUNK = 16;

template <int LT>
__global__ void __launch_bounds__(256) ks_inner_kernel(const KsParams P, const ModConst *__restrict__ mods)
{
    const u32 L = LT ? LT : P.L, n = P.n;
    const u32 x = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 g = blockIdx.y / (L + 1), i = blockIdx.y % (L + 1);
    const u32 b0 = blockIdx.z * KS_INNER_BCHUNK;
    const u32 b1 = min(P.B, b0 + KS_INNER_BCHUNK);
    if (x >= n) return;
    const u32 ki = (i == L) ? P.K - 1 : i;
    const ModConst m = mods[ki];
    const u64 *key = P.key[g] + (size_t)ki * n + x;
    const size_t kstride = (size_t)P.K * n;  // between (j,c) slices
    u64 k0[LT ? LT : 1], k1[LT ? LT : 1];
    if (LT) {
#pragma unroll
        for (int j = 0; j < (LT ? LT : 1); ++j) {
            k0[j] = __ldg(key + (size_t)(2 * j) * kstride);
            k1[j] = __ldg(key + (size_t)(2 * j + 1) * kstride);
        }
    }
    const CtView &v = P.in[g];
    const u32 *pm = P.perm[g];
    const u32 xs = ((i < L || P.hoisted) && pm) ? __ldg(pm + x) : x;  // gathered column
    const u32 xe = P.hoisted ? xs : x;                                  // column of the lifted digits
    for (u32 b = b0; b < b1; ++b) {
        const size_t e = (size_t)g * P.B + b;
        const size_t ee = P.hoisted ? b : e;
        u64 h0 = 0, l0 = 0, h1 = 0, l1 = 0;
        if (LT) {
            u64 d[LT ? LT : 1];
#pragma unroll
            for (int j = 0; j < (LT ? LT : 1); ++j)
                d[j] = ((u32)j == i) ? v.p[b * v.sb + P.target_poly * v.sp + j * v.sl + xs]
                                     : P.ext[((ee * L + j) * (L + 1) + i) * n + xe];
#pragma unroll
            for (int j = 0; j < (LT ? LT : 1); ++j) {
                mac128(h0, l0, d[j], k0[j]);
                mac128(h1, l1, d[j], k1[j]);
            }
        } else {
            for (u32 j = 0; j < L; ++j) {
                const u64 d = (j == i) ? v.p[b * v.sb + P.target_poly * v.sp + j * v.sl + xs]
                                       : P.ext[((ee * L + j) * (L + 1) + i) * n + xe];
                mac128(h0, l0, d, __ldg(key + (size_t)(2 * j) * kstride));
                mac128(h1, l1, d, __ldg(key + (size_t)(2 * j + 1) * kstride));
            }
        }
        if (P.add_pc0 && i < L) mac128(h0, l0, v.p[b * v.sb + i * v.sl + xs], m.pmont);
        P.acc[((e * 2 + 0) * (L + 1) + i) * n + x] = mont_reduce(h0, l0, m);  // keys are stored as k*2^64 mod m
        P.acc[((e * 2 + 1) * (L + 1) + i) * n + x] = mont_reduce(h1, l1, m);
    }
}

// ---------------------------------------------------------------------------------------
// element-wise kernels over (b, p, l, x).  b's batch stride may be 0 (broadcast).
// ---------------------------------------------------------------------------------------
enum EwOp { EW_ADD, EW_SUB, EW_NEG, EW_COPY, EW_MULPLAIN, EW_ADDPLAIN, EW_SUBPLAIN, EW_NEGCOPY_B };

struct EwParams {
    CtView out, a, b;  // b: second ciphertext, or plaintext (sp = 0)
    u32 B, polys, L, n;
};

template <int OP>
__global__ void __launch_bounds__(256) ew_kernel(const EwParams P, const ModConst *__restrict__ mods)
{
    const size_t per_b = (size_t)P.polys * P.L * P.n;
    const size_t total = (size_t)P.B * per_b;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const u32 b = (u32)(idx / per_b);
        u32 r = (u32)(idx % per_b);
        const u32 x = r % P.n;
        r /= P.n;
        const u32 l = r % P.L, p = r / P.L;
        const size_t oa = b * P.a.sb + p * P.a.sp + l * P.a.sl + x;
        const size_t ob = b * P.b.sb + p * P.b.sp + l * P.b.sl + x;
        const size_t oo = b * P.out.sb + p * P.out.sp + l * P.out.sl + x;
        const u64 q = mods[l].q;
        u64 v;
        if (OP == EW_ADD) v = addmod(P.a.p[oa], P.b.p[ob], q);
        else if (OP == EW_SUB) v = submod(P.a.p[oa], P.b.p[ob], q);
        else if (OP == EW_NEG) v = negmod(P.a.p[oa], q);
        else if (OP == EW_COPY) v = P.a.p[oa];
        else if (OP == EW_NEGCOPY_B) v = negmod(P.b.p[ob], q);
        else if (OP == EW_MULPLAIN) v = mulmod(P.a.p[oa], P.b.p[ob], mods[l]);
        else if (OP == EW_ADDPLAIN) v = addmod(P.a.p[oa], P.b.p[ob], q);
        else v = submod(P.a.p[oa], P.b.p[ob], q);
        P.out.p[oo] = v;
    }
}

// K5: tensor product 2x2 -> 3 (d0 = a0 b0, d1 = a0 b1 + a1 b0, d2 = a1 b1); SQUARE: b = a.
// out may alias a or b: every thread reads its inputs before it writes.
template <bool SQUARE>
__global__ void __launch_bounds__(256) tensor_kernel(const EwParams P, const ModConst *__restrict__ mods)
{
    const size_t per_b = (size_t)P.L * P.n;
    const size_t total = (size_t)P.B * per_b;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const u32 b = (u32)(idx / per_b);
        const u32 r = (u32)(idx % per_b);
        const u32 l = r / P.n, x = r % P.n;
        const ModConst m = mods[l];
        const size_t oa = b * P.a.sb + l * P.a.sl + x;
        const u64 a0 = P.a.p[oa], a1 = P.a.p[oa + P.a.sp];
        u64 b0, b1;
        if (SQUARE) {
            b0 = a0;
            b1 = a1;
        } else {
            const size_t ob = b * P.b.sb + l * P.b.sl + x;
            b0 = P.b.p[ob];
            b1 = P.b.p[ob + P.b.sp];
        }
        const u64 d0 = mulmod(a0, b0, m), d2 = mulmod(a1, b1, m);
        u64 h = 0, lo = 0;
        mac128(h, lo, a0, b1);
        mac128(h, lo, a1, b0);
        const u64 d1 = barrett128(h, lo, m);
        const size_t oo = b * P.out.sb + l * P.out.sl + x;
        P.out.p[oo] = d0;
        P.out.p[oo + P.out.sp] = d1;
        P.out.p[oo + 2 * P.out.sp] = d2;
    }
}

// general ciphertext product (sizes sa x sb -> sa+sb-1), used when an operand has size 3.
// out must not alias.
__global__ void __launch_bounds__(256) convolve_kernel(const EwParams P, u32 sa, u32 sb, const ModConst *__restrict__ mods)
{
    const u32 so = sa + sb - 1;
    const size_t per_b = (size_t)so * P.L * P.n;
    const size_t total = (size_t)P.B * per_b;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const u32 b = (u32)(idx / per_b);
        u32 r = (u32)(idx % per_b);
        const u32 x = r % P.n;
        r /= P.n;
        const u32 l = r % P.L, k = r / P.L;
        const ModConst m = mods[l];
        u64 h = 0, lo = 0;
        for (u32 ia = 0; ia < sa; ++ia) {
            if (k < ia || k - ia >= sb) continue;
            mac128(h, lo, P.a.p[b * P.a.sb + ia * P.a.sp + l * P.a.sl + x], P.b.p[b * P.b.sb + (k - ia) * P.b.sp + l * P.b.sl + x]);
        }
        P.out.p[b * P.out.sb + k * P.out.sp + l * P.out.sl + x] = barrett128(h, lo, m);
    }
}

// K6 standalone (only used when a rotation's output aliases its input): out = pi(in)
__global__ void __launch_bounds__(256) fixup_kernel(const CtView v, u32 B, u32 polys, u32 L, u32 n, const ModConst *__restrict__ mods)
{
    const size_t per_b = (size_t)polys * L * n;
    const size_t total = (size_t)B * per_b;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const u32 b = (u32)(idx / per_b);
        u32 r = (u32)(idx % per_b);
        const u32 x = r % n;
        r /= n;
        const u32 l = r % L, p = r / L;
        u64 *ptr = v.p + b * v.sb + p * v.sp + l * v.sl + x;
        *ptr = barrett64(*ptr, mods[l]);
    }
}

// ---------------------------------------------------------------------------------------
// K10: BSGS inner sums for every giant step in one pass over the baby rotations.
//   inner[g][b][p][l][x] = sum_{k<n1} baby_k[b][p][l][x] * diag[g*n1+k][l][x]  mod q_l
// ---------------------------------------------------------------------------------------
struct BsgsParams {
    CtView baby[MAXB];  // baby[0] = the input batch
    CtView inner;       // batch index = g*B + b
    const u64 *diag;    // [n1*n2][Lcap][N], Montgomery form (d * 2^64 mod q_l)
    size_t diag_si, diag_sl;
    u32 n1, n2, B, L, n;
    u32 special_limb, K;  // limb index that lives mod the special prime (mods[K-1]); ~0u: none
};

// block = (32 coefficients) x (BSGS_GT giant steps); grid = (n/32, L, ceil(n2/BSGS_GT)).
// A thread keeps the N1 diagonal words of its (g, l, x) in registers for the whole batch and
// sweeps the 2B polynomials.  The rotated-ciphertext words of BSGS_C consecutive polynomials are
// staged in shared memory by cp.async (double buffered) and shared by the BSGS_GT warps, so
// they are read from HBM exactly once; the diagonals are read exactly once as well.
constexpr int BSGS_GT = 8;

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int N1>
__global__ void __launch_bounds__(32 * BSGS_GT) bsgs_inner_kernel(const BsgsParams P, const ModConst *__restrict__ mods)
{
    constexpr int BSGS_C = N1 > 16 ? 2 : 4;  // staging depth: 2 x C x N1 x 256 B of static shared memory
    __shared__ u64 buf[2][BSGS_C][N1][32];
    __shared__ const u64 *bptr[N1];
    __shared__ size_t bsb[N1], bsp[N1];
    const u32 tid = threadIdx.y * 32 + threadIdx.x;
    const u32 x0 = blockIdx.x * 32, l = blockIdx.y;
    const u32 g = blockIdx.z * BSGS_GT + threadIdx.y;
    const bool active = g < P.n2;
    if (tid < N1) {
        const CtView &c = P.baby[tid];
        bptr[tid] = c.p + l * c.sl + x0;
        bsb[tid] = c.sb;
        bsp[tid] = c.sp;
    }
    __syncthreads();
    const ModConst m = mods[l == P.special_limb ? P.K - 1 : l];
    u64 d[N1];
    if (active) {
        const u64 *dp = P.diag + (size_t)g * N1 * P.diag_si + l * P.diag_sl + x0 + threadIdx.x;
#pragma unroll
        for (int k = 0; k < N1; ++k) d[k] = __ldg(dp + k * P.diag_si);
    }
    const u32 iters = 2 * P.B;
    const u32 nchunks = (iters + BSGS_C - 1) / BSGS_C;
    auto issue = [&](u32 chunk, u32 s) {
        for (u32 v = tid; v < BSGS_C * N1 * 32; v += 32 * BSGS_GT) {
            const u32 xx = v & 31, k = (v >> 5) % N1, ci = v / (32 * N1);
            const u32 it = chunk * BSGS_C + ci;
            if (it < iters) cp_async8(&buf[s][ci][k][xx], bptr[k] + (it >> 1) * bsb[k] + (it & 1) * bsp[k] + xx);
        }
        cp_async_commit();
    };
    issue(0, 0);
    for (u32 ch = 0; ch < nchunks; ++ch) {
        const u32 s = ch & 1;
        if (ch + 1 < nchunks) {
            issue(ch + 1, s ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (active) {
#pragma unroll
            for (int ci = 0; ci < BSGS_C; ++ci) {
                const u32 it = ch * BSGS_C + ci;
                if (it < iters) {
                    u64 h = 0, lo = 0;
#pragma unroll
                    for (int k = 0; k < N1; ++k) mac128(h, lo, buf[s][ci][k][threadIdx.x], d[k]);
                    P.inner.p[((size_t)g * P.B + (it >> 1)) * P.inner.sb + (it & 1) * P.inner.sp + l * P.inner.sl + x0 + threadIdx.x] =
                        N1 > 16 ? mont_reduce_wide(h, lo, m) : mont_reduce(h, lo, m);  // diag holds d*2^64 mod q
                }
            }
        }
        __syncthreads();  // the buffer is refilled by the next-but-one issue
    }
}

// ---------------------------------------------------------------------------------------
// Double-hoisted BSGS, fused: baby-step key inner products + inner sums of every giant step in one
// pass (replaces scale_by_p + ks_inner(add_pc0) + bsgs_inner of the unfused path; same values):
//   b_0[c]   = P * ct[c]                                              (limb L: 0)
//   b_k[c]   = sum_j pi_k(digit_j) * key_k[j][c]  (+ P * pi_k(c0) for c = 0, limbs < L)
//   u_g[c]   = sum_k b_k[c] * diag[g*n1+k]                            all in the extended basis
// A CTA owns one extended limb i and a tile of DH_TX = 32 coefficients; it stages the key words
// (n1 x 2L), the diagonal words (n1*n2) and the gather indices of that tile in shared memory ONCE
// and sweeps a chunk of the batch, so keys and diagonals leave HBM once per chunk and the rotated
// ciphertexts b_k never exist in HBM.  Inside the CTA every WARP owns whole ciphertexts (lane =
// coefficient): it walks the n1 baby steps, builds b_k in registers (the lifted digits are gathered
// from HBM/L2 -- a Galois permutation maps an aligned 32-word tile to an aligned tile, so a gather
// is one fully used 256-byte line -- and prefetched two baby steps ahead) and feeds it straight
// into the 2*n2 independent 128-bit accumulators of the giant steps: no barrier and no
// shared-memory round trip in the main loop, one Montgomery reduction per output word.
// grid = (N / 32, L+1, batch chunks), block = 32 * DH_KG.
// ---------------------------------------------------------------------------------------
constexpr int DH_TX = 32, DH_KG = 8, DH_BCH = 32;
struct DhInnerParams {
    CtView in;             // input batch at level L
    const u64 *ext;        // [B][L][L+1][N] lifted digits of c1 (hoisted decomposition)
    const u64 *key[MAXB];  // Galois key of baby step k, Montgomery form; [0] unused
    const u32 *perm[MAXB]; // gather table of baby step k; [0] unused
    const u64 *diag;       // [n1*n2][Lcap][N] Montgomery form, limb L = special prime
    size_t diag_si;
    u64 *u;                // [n2][B][2][L+1][N]
    u32 n1, n2, B, L, K, n;
};
static inline size_t dh_inner_smem(u32 n1, u32 n2, u32 L)
{
    return ((size_t)n1 * 2 * L + (size_t)n1 * n2) * DH_TX * sizeof(u64) + (size_t)n1 * DH_TX * sizeof(u32);
}

template <int LT, int N2>
__global__ void __launch_bounds__(DH_TX * DH_KG, 2) dh_inner_kernel(const DhInnerParams P, const ModConst *__restrict__ mods)
{
    extern __shared__ __align__(16) u64 dh_smem[];
    constexpr u32 TX = DH_TX, NT = DH_TX * DH_KG, L = LT;
    const u32 n = P.n, n1 = P.n1;
    const u32 lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const u32 x0 = blockIdx.x * TX, i = blockIdx.y;
    const u32 b0 = blockIdx.z * DH_BCH, b1 = min(P.B, b0 + DH_BCH);
    const u32 ki = (i == L) ? P.K - 1 : i;
    const ModConst m = mods[ki];
    u64 *skey = dh_smem;                                        // [n1][2L][TX]
    u64 *sdiag = skey + (size_t)n1 * 2 * L * TX;                 // [n1][N2][TX]
    u32 *sperm = reinterpret_cast<u32 *>(sdiag + (size_t)n1 * N2 * TX);  // [n1][TX]
    for (u32 idx = threadIdx.x; idx < n1 * 2 * L * TX; idx += NT) {
        const u32 x = idx % TX, r = idx / TX, jc = r % (2 * L), k = r / (2 * L);
        skey[idx] = k ? __ldg(P.key[k] + ((size_t)jc * P.K + ki) * n + x0 + x) : 0;
    }
    for (u32 idx = threadIdx.x; idx < n1 * N2 * TX; idx += NT) {
        const u32 x = idx % TX, r = idx / TX, g = r % N2, k = r / N2;
        sdiag[idx] = g < P.n2 ? __ldg(P.diag + (size_t)(g * n1 + k) * P.diag_si + (size_t)i * n + x0 + x) : 0;
    }
    for (u32 idx = threadIdx.x; idx < n1 * TX; idx += NT) {
        const u32 x = idx % TX, k = idx / TX;
        sperm[idx] = k ? __ldg(P.perm[k] + x0 + x) : x0 + x;
    }
    __syncthreads();
    const bool data_limb = i < L;
    for (u32 b = b0 + w; b < b1; b += DH_KG) {
        const u64 *cb = P.in.p + b * P.in.sb;
        const u64 *eb = P.ext + ((size_t)b * L * (L + 1) + i) * n;
        // operands of baby step k: k = 0 -> {c0, c1} of the CTA's own limb; k >= 1 -> the L permuted
        // digits (digit i itself is c1's limb i) and, on data limbs, the permuted c0 word
        auto fetch = [&](u32 k, u64 (&d)[LT + 1]) {
            const u32 xs = sperm[k * TX + lane];
            if (k == 0) {
                d[0] = data_limb ? cb[i * P.in.sl + xs] : 0;
                d[1] = data_limb ? cb[P.in.sp + i * P.in.sl + xs] : 0;
#pragma unroll
                for (int j = 2; j <= LT; ++j) d[j] = 0;
                return;
            }
#pragma unroll
            for (int j = 0; j < LT; ++j)
                d[j] = ((u32)j == i) ? cb[P.in.sp + j * P.in.sl + xs] : eb[(size_t)j * (L + 1) * n + xs];
            d[LT] = data_limb ? cb[i * P.in.sl + xs] : 0;
        };
        u64 ah[N2][2], al[N2][2];
#pragma unroll
        for (int g = 0; g < N2; ++g) ah[g][0] = al[g][0] = ah[g][1] = al[g][1] = 0;
        u64 d0[LT + 1], d1[LT + 1], d2[LT + 1];
        fetch(0, d0);
        if (n1 > 1) fetch(1, d1);
        // one baby step: start the gathers of step k+2 into `nxt`, then consume `cur`
        auto step = [&](u32 k, const u64 (&cur)[LT + 1], u64 (&nxt)[LT + 1]) {
            if (k + 2 < n1) fetch(k + 2, nxt);
            u64 h0 = 0, l0 = 0, h1 = 0, l1 = 0;
            if (k == 0) {
                mac128(h0, l0, cur[0], m.pmont);  // pmont = 0 for the special prime: b_0[L] = 0
                mac128(h1, l1, cur[1 % (LT + 1)], m.pmont);
            } else {
                const u64 *kp = skey + (size_t)k * 2 * L * TX + lane;
#pragma unroll
                for (int j = 0; j < LT; ++j) {
                    mac128(h0, l0, cur[j], kp[(2 * j) * TX]);
                    mac128(h1, l1, cur[j], kp[(2 * j + 1) * TX]);
                }
                mac128(h0, l0, cur[LT], m.pmont);
            }
            const u64 a0 = mont_reduce(h0, l0, m), a1 = mont_reduce(h1, l1, m);
            const u64 *dp = sdiag + (size_t)k * N2 * TX + lane;
#pragma unroll
            for (int g = 0; g < N2; ++g) {
                const u64 dg = dp[g * TX];
                mac128(ah[g][0], al[g][0], a0, dg);
                mac128(ah[g][1], al[g][1], a1, dg);
            }
        };
        for (u32 k = 0; k < n1; k += 3) {  // the three operand sets rotate by name, not by copying
            step(k, d0, d2);
            if (k + 1 < n1) step(k + 1, d1, d0);
            if (k + 2 < n1) step(k + 2, d2, d1);
        }
#pragma unroll
        for (int g = 0; g < N2; ++g) {
            if ((u32)g < P.n2) {
                u64 *up = P.u + ((((size_t)g * P.B + b) * 2) * (L + 1) + i) * n + x0 + lane;
                up[0] = mont_reduce_wide(ah[g][0], al[g][0], m);
                up[(size_t)(L + 1) * n] = mont_reduce_wide(ah[g][1], al[g][1], m);
            }
        }
    }
}

// sum of `terms` ciphertext batches laid out at batch offsets g*B: out = sum_g src_g
struct SumParams {
    CtView first;  // term 0
    CtView rest;   // terms 1.., batch index (g-1)*B + b
    CtView out;
    u32 terms, B, polys, L, n;
};
__global__ void __launch_bounds__(256) sum_terms_kernel(const SumParams P, const ModConst *__restrict__ mods)
{
    const size_t per_b = (size_t)P.polys * P.L * P.n;
    const size_t total = (size_t)P.B * per_b;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const u32 b = (u32)(idx / per_b);
        u32 r = (u32)(idx % per_b);
        const u32 x = r % P.n;
        r /= P.n;
        const u32 l = r % P.L, p = r / P.L;
        const u64 q = mods[l].q;
        u64 s = P.first.p[b * P.first.sb + p * P.first.sp + l * P.first.sl + x];
        for (u32 g = 1; g < P.terms; ++g)
            s = addmod(s, P.rest.p[((size_t)(g - 1) * P.B + b) * P.rest.sb + p * P.rest.sp + l * P.rest.sl + x], q);
        P.out.p[b * P.out.sb + p * P.out.sp + l * P.out.sl + x] = s;
    }
}

// lazy mod-down over giant steps: accsum[b][c][i] = sum_g acc[g*B+b][c][i] mod m_i  (rows = 2*(L+1))
__global__ void __launch_bounds__(256) acc_group_sum_kernel(const u64 *__restrict__ acc, const u64 *__restrict__ init,
                                                            u64 *__restrict__ out, u32 groups, u32 B, u32 L, u32 K, u32 n,
                                                            const ModConst *__restrict__ mods)
{
    const size_t per_b = (size_t)2 * (L + 1) * n;
    const size_t total = (size_t)B * per_b;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const u32 b = (u32)(idx / per_b);
        const size_t r = idx % per_b;
        const u32 i = (u32)((r / n) % (L + 1));
        const u64 q = mods[i == L ? K - 1 : i].q;
        u64 s = init ? init[idx] : 0;  // init: an unrotated term already in the extended basis
        for (u32 g = 0; g < groups; ++g) s = addmod(s, acc[((size_t)g * B + b) * per_b + r], q);
        out[idx] = s;
    }
}

// base[b][0] = first[b][0] + sum_g pi_g(rest[g*B+b][0]);  base[b][1] = first[b][1]
struct BaseSumParams {
    CtView first, rest, out;
    const u32 *perm[MAXG];
    u32 groups, B, L, n;
    u32 has_first;  // 0: there is no unrotated term (diagonal-sharded ranks other than the first)
};
__global__ void __launch_bounds__(256) base_gather_sum_kernel(const BaseSumParams P, const ModConst *__restrict__ mods)
{
    const size_t per_b = (size_t)2 * P.L * P.n;
    const size_t total = (size_t)P.B * per_b;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const u32 b = (u32)(idx / per_b);
        u32 r = (u32)(idx % per_b);
        const u32 x = r % P.n;
        r /= P.n;
        const u32 l = r % P.L, p = r / P.L;
        const u64 q = mods[l].q;
        u64 s = P.has_first ? P.first.p[b * P.first.sb + p * P.first.sp + l * P.first.sl + x] : 0;
        if (p == 0)
            for (u32 g = 0; g < P.groups; ++g)
                s = addmod(s, P.rest.p[((size_t)g * P.B + b) * P.rest.sb + l * P.rest.sl + __ldg(P.perm[g] + x)], q);
        P.out.p[b * P.out.sb + p * P.out.sp + l * P.out.sl + x] = s;
    }
}

// x -> x * 2^64 mod q (Montgomery form) for `rows` limb polynomials; row r uses modulus
// mod_of_row = (r % limbs_per_item) mapped through `last_is_special` (key layout: K limbs, limb
// K-1 = special prime; plaintext layout: limb l = modulus l)
__global__ void __launch_bounds__(256) to_montgomery_kernel(const u64 *__restrict__ src, u64 *__restrict__ dst, size_t rows,
                                                            u32 limbs_per_item, size_t row_stride_src, size_t row_stride_dst, u32 n,
                                                            const ModConst *__restrict__ mods, u32 special_limb = ~0u, u32 K = 0)
{
    const size_t total = rows * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t r = idx / n;
        const u32 x = (u32)(idx % n);
        const u32 l = (u32)(r % limbs_per_item);
        const ModConst m = mods[l == special_limb ? K - 1 : l];
        dst[r * row_stride_dst + x] = mul_shoup(src[r * row_stride_src + x], m.rmod, m.rmod_sh, m.q);
    }
}

// double-hoisted baby step 0: out[b][c][i] = P * in[b][c][i] mod q_i (i < L), out[b][c][L] = 0;
// out in the key-switch accumulator layout [B][2][L+1][N]
__global__ void __launch_bounds__(256) scale_by_p_kernel(const CtView in, u64 *__restrict__ out, u32 B, u32 L, u32 n,
                                                         const ModConst *__restrict__ mods)
{
    const size_t per_b = (size_t)2 * (L + 1) * n;
    const size_t total = (size_t)B * per_b;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const u32 b = (u32)(idx / per_b);
        u32 r = (u32)(idx % per_b);
        const u32 x = r % n;
        r /= n;
        const u32 i = r % (L + 1), c = r / (L + 1);
        u64 v = 0;
        if (i < L) {
            const ModConst m = mods[i];
            u64 h = 0, lo = 0;
            mac128(h, lo, in.p[b * in.sb + c * in.sp + i * in.sl + x], m.pmont);
            v = mont_reduce(h, lo, m);
        }
        out[idx] = v;
    }
}

// strided repack between SEAL's packed host layout (staged in HBM) and the batch layout
__global__ void __launch_bounds__(256) repack_kernel(const CtView dst, const CtView src, u32 B, u32 polys, u32 L, u32 n)
{
    const size_t per_b = (size_t)polys * L * n;
    const size_t total = (size_t)B * per_b;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const u32 b = (u32)(idx / per_b);
        u32 r = (u32)(idx % per_b);
        const u32 x = r % n;
        r /= n;
        const u32 l = r % L, p = r / L;
        dst.p[b * dst.sb + p * dst.sp + l * dst.sl + x] = src.p[b * src.sb + p * src.sp + l * src.sl + x];
    }
}

}  // namespace hegpu
