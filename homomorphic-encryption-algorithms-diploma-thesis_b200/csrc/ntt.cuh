// ntt.cuh -- single-pass negacyclic NTT / INTT of one RNS limb polynomial per CTA.
//
// Replaces seal::util::ntt_negacyclic_harvey / inverse_ntt_negacyclic_harvey (SURVEY.md
// 2.1 K1/K2, spec 9.2): forward = Cooley-Tukey, natural-order input -> bit-reversed-order
// output; inverse = Gentleman-Sande, exact inverse with N^-1 folded into the last stage.
//
// B200 mapping (DESIGN.md "NTT"):
//  * one CTA owns one limb polynomial (N <= 16384 words = 128 KiB of shared memory), so a
//    transform moves the algorithmic 16*N bytes through HBM exactly once;
//  * every thread keeps 16 coefficients in registers and runs a radix-16 pass (4 butterfly
//    stages) on them (512 threads x 2 sets at N = 16384, so 128 registers per thread); passes are separated by an in-place exchange through shared memory
//    with an XOR swizzle that keeps every 64-bit access bank-conflict free;
//  * the first pass reads global memory directly (coalesced, through a caller-supplied
//    loader that can gather / lift / subtract) and the last pass stores through a
//    caller-supplied epilogue, so key-switch and rescale fuse their element-wise steps
//    into the transforms;
//  * three arithmetic policies, chosen per modulus (uniform per CTA):
//      ArI64<true>   60-bit primes: Shoup multiplication (6 IMAD.WIDE + 4 IMAD per butterfly), lazy
//                    ranges with one correction per radix-16 pass (fwd) / per stage (inv);
//      ArI64<false>  primes below 2^58 (fwd) / 2^46 (inv): same, no corrections at all;
//      ArF64         primes below 2^43: the butterfly runs on the FP64 pipe (64 FMA/clk/SM on
//                    B200) -- exact integer arithmetic in doubles: h = y*w, l = fma(y,w,-h),
//                    t = rint(h/q), r = fma(-t,q,h) + l == y*w - t*q exactly; 8 FP64 ops per
//                    butterfly instead of ~41 integer-multiply issue cycles;
//  * N = 16384 runs as ONE CTA that transforms the two halves one after the other in 64 KiB of
//    shared memory ("park" kernels): the stride-N/2 stage is done on the fly while loading, the
//    half that is not being worked on is parked in an L2-resident scratch (+8*N bytes of L2
//    traffic, no extra HBM pass), and 2-3 such CTAs share an SM so that their load, exchange and
//    store phases overlap (measured +35 % over one 128 KiB CTA per SM);
//  * N = 32768 runs as two CTAs that each redo the first (stride N/2) stage from global
//    memory and then own one half (forward), or finish with one element-wise stage kernel
//    (inverse) -- see ntt_fwd_kernel / ntt_inv_kernel SPLIT.
#pragma once
#include <type_traits>

#include "modarith.cuh"

namespace hegpu {

// double-precision constants of a modulus q < 2^43 (FP64 butterfly path)
struct ModF64 {
    double q, qinv, ninv, wlast;  // wlast = inv[1]*N^-1 mod q
};

struct NttTables {
    const ulonglong2 *fwd;  // [K][N]  {psi^brev(i), shoup}   index m+i as in SURVEY 9.2
    const ulonglong2 *inv;  // [K][N]  {psi^-brev(i), shoup}
    const ModConst *mods;   // [K]
    const ulonglong2 *inv_last;  // [K] {inv[1]*N^-1, shoup}: bottom twiddle of the last INTT stage
    const double *fwd_d;    // [K][N] the same twiddles as doubles (valid where mods[i].big & 4)
    const double *inv_d;
    const ModF64 *modsd;    // [K]
    u32 n;                  // ring degree N
    u32 logn;
};

// LOGE = log2 of the coefficients a thread keeps in registers per set: 4 (radix-16 passes, 512
// threads x 128 registers) or 3 (radix-8 passes, 1024 threads x 64 registers: twice the warps).
template <int LOGL, int LOGE>
struct NttShape {
    static constexpr int LSIZE = 1 << LOGL;
    static constexpr int E = 1 << LOGE;
    static constexpr int SETS = LSIZE / E;                        // register sets per pass
    // <= 64 KiB transforms run 256 threads so that 2 (LOGE = 4, 128 registers) or 3 (LOGE = 3, 64
    // registers) CTAs share an SM and overlap their load / exchange / store phases
    static constexpr int MAXT = (LOGL <= 13) ? 256 : ((LOGE == 4) ? 512 : 1024);
#ifndef HEGPU_MINB_LOGE3
#define HEGPU_MINB_LOGE3 3
#endif
    static constexpr int MINB = (LOGL <= 13) ? ((LOGE == 4) ? 2 : HEGPU_MINB_LOGE3) : 1;
    static constexpr int THREADS = SETS > MAXT ? MAXT : SETS;
    static constexpr int ITER = SETS / THREADS;
    static constexpr int REM = (LOGL % LOGE == 0) ? LOGE : (LOGL % LOGE);  // stages of the partial pass
    static constexpr int NFULL = (LOGL - REM) / LOGE;                      // number of full passes
    static constexpr size_t SMEM = sizeof(u64) << LOGL;
};

// Bank swizzle of the in-place exchange buffer (64-bit words, 16 banks per half-warp access): two ALU operations per
// address.  The pass whose register field sits at index bit 1 runs with 2-way conflicts under this fold (index bits 0
// and 4 land in the same bank bit; ncu: 23 % of a transform's shared-memory wavefronts).  The conflict-free linear map
// idx[3:0] ^ (h ^ h << 1), h = idx[7:4], was built and measured: bit-identical, zero conflicts, and 1-2 % SLOWER --
// it costs two more ALU operations per address in kernels that are bound by instruction issue, not by shared memory.
__device__ __forceinline__ u32 swz(u32 idx) { return idx ^ ((idx >> 4) & 15u); }

// Position of the 2^LOGE coefficients a thread owns in a pass.  The low S bits of the register
// index k sit at index bits [PLO+S-1:PLO]; the remaining LOGE-S bits sit at [PHI+LOGE-1-S:PHI];
// the thread id fills every other bit, low to high.
template <int LOGL, int LOGE, int S, int PLO, int PHI>
struct PassMap {
    __device__ static __forceinline__ u32 base(u32 t)
    {
        u32 lo = t & ((1u << PLO) - 1u);
        u32 mid = (t >> PLO) & ((1u << (PHI - PLO - S)) - 1u);
        u32 hi = (S == LOGE) ? 0u : (t >> (PHI - S));
        u32 r = lo | (mid << (PLO + S));
        if (S != LOGE) r |= hi << (PHI + LOGE - S);
        return r;
    }
    __device__ static __forceinline__ u32 off(int k)
    {
        u32 r = (u32)(k & ((1 << S) - 1)) << PLO;
        if (S != LOGE) r |= (u32)(k >> S) << PHI;
        return r;
    }
};

// ------------------------------------------------------------------ arithmetic policies
#define HEGPU_TWO52 4503599627370496.0
#define HEGPU_MAGIC 6755399441055744.0 /* 1.5 * 2^52: rint() by addition for |x| < 2^51 */

#ifndef HEGPU_SHOUP_APPROX
#define HEGPU_SHOUP_APPROX 1  // 0: every butterfly uses the exact Shoup quotient (the round-1 form)
#endif
template <bool BIG>
struct ArI64 {
    typedef u64 V;
    typedef ulonglong2 TW;
    const ModConst &m;
    const ulonglong2 wlast;
    const u64 nq;  // 2^64 - q
    const u64 q3;  // 3q: lazy bound of a product with the approximate quotient
    __device__ __forceinline__ ArI64(const ModConst &m_, ulonglong2 wl) : m(m_), wlast(wl), nq(0ull - m_.q), q3(m_.q3) {}
    __device__ __forceinline__ V from_load(u64 v) const { return v; }
    // Forward lazy ranges.  BIG (q < 2^60, so 16q < 2^64): values < 8q at pass start, a stage adds 2q with the exact
    // quotient (T < 2q) and 3q with the approximate one (T < 3q); a pass of S stages may use the approximate form in
    // approx_stages(S) = min(S, 8 - 2S) of them and still end below 16q (radix-8 passes: two of three stages).
    // !BIG (q < 2^58): < 4q at the start (fold loaders), 14 more stages of at most 3q stay below 46q < 2^64: always approximate.
    static __host__ __device__ constexpr int approx_stages(int S)
    {
        if (!HEGPU_SHOUP_APPROX) return 0;
        if (!BIG) return S;
        return 8 - 2 * S < 0 ? 0 : (8 - 2 * S < S ? 8 - 2 * S : S);
    }
    __device__ __forceinline__ V fwd_fix(V v) const { return BIG ? csub(v, m.q << 3) : v; }
    template <bool APPROX>
    __device__ __forceinline__ void fwd_bfly(V &X, V &Y, const TW W) const
    {
        if constexpr (APPROX) {
            const u64 T = mul_shoup_lazy3_nq(Y, W.x, W.y, nq);
            Y = X + q3 - T;
            X = X + T;
        } else {
            const u64 T = mul_shoup_lazy_nq(Y, W.x, W.y, nq);
            Y = X + (m.q << 1) - T;
            X = X + T;
        }
    }
    __device__ __forceinline__ u64 fwd_final(V v) const
    {
        if (BIG) {
            v = csub(v, m.q << 3);
            v = csub(v, m.q << 2);
            v = csub(v, m.q << 1);
            return csub(v, m.q);
        }
        return barrett64(v, m);
    }
    // Inverse lazy ranges: products come from the approximate quotient and lie in [0, LZ q), LZ = 3 (2 with the exact one).
    // BIG keeps [0, LZ q) with a correction of the sum per stage; !BIG (q < 2^46) lets the sums double (bound
    // LZ q * 2^gbit entering global stage gbit; the difference is offset by 4q * 2^gbit).
    static constexpr int LZ = HEGPU_SHOUP_APPROX ? 3 : 2;
    __device__ __forceinline__ u64 lzq() const { return HEGPU_SHOUP_APPROX ? q3 : m.q << 1; }
    __device__ __forceinline__ u64 mul_lazy(u64 x, const TW W) const
    {
        return HEGPU_SHOUP_APPROX ? mul_shoup_lazy3_nq(x, W.x, W.y, nq) : mul_shoup_lazy_nq(x, W.x, W.y, nq);
    }
    __device__ __forceinline__ V inv_fix(V v) const { return v; }
    template <int GBIT>
    __device__ __forceinline__ void inv_bfly(V &X, V &Y, const TW W) const
    {
        u64 Sm, D;
        if (BIG) {
            Sm = csub(X + Y, lzq());
            D = X + lzq() - Y;
        } else {
            Sm = X + Y;
            D = X + (m.q << (GBIT + 2)) - Y;
        }
        X = Sm;
        Y = mul_lazy(D, W);
    }
    template <int GBIT>
    __device__ __forceinline__ void inv_bfly_last(V &X, V &Y) const
    {
        // the exact products accept any 64-bit operand: the sum needs no correction here
        const u64 Sm = X + Y;
        const u64 D = BIG ? X + lzq() - Y : X + (m.q << (GBIT + 2)) - Y;
        X = mul_shoup(Sm, m.ninv, m.ninv_sh, m.q);
        Y = mul_shoup(D, wlast.x, wlast.y, m.q);
    }
    __device__ __forceinline__ u64 inv_final(V v) const { return v; }
};

struct ArF64 {
    typedef double V;
    typedef double TW;
    const ModF64 f;
    __device__ __forceinline__ explicit ArF64(const ModF64 &f_) : f(f_) {}
    // exact y*w - t*q with t = rint(y*w/q): |result| <= ~0.7q for |y| < 2^48, 0 <= w < q < 2^43
    __device__ __forceinline__ double mulmod(double y, double w) const
    {
        const double h = __dmul_rn(y, w);
        const double l = __fma_rn(y, w, -h);
        const double t = __dadd_rn(__fma_rn(h, f.qinv, HEGPU_MAGIC), -HEGPU_MAGIC);
        return __dadd_rn(__fma_rn(-t, f.q, h), l);
    }
    __device__ __forceinline__ double reduce(double v) const
    {
        const double t = __dadd_rn(__fma_rn(v, f.qinv, HEGPU_MAGIC), -HEGPU_MAGIC);
        return __fma_rn(-t, f.q, v);
    }
    __device__ __forceinline__ u64 to_canonical(double r) const  // |r| < q
    {
        r = r < 0.0 ? __dadd_rn(r, f.q) : r;
        return (u64)__double_as_longlong(__dadd_rn(r, HEGPU_TWO52)) & 0xFFFFFFFFFFFFFull;
    }
    __device__ __forceinline__ V from_load(u64 v) const  // v < 2^52
    {
        return __dadd_rn(__longlong_as_double((long long)(v | 0x4330000000000000ull)), -HEGPU_TWO52);
    }
    // forward: |T| <= 0.7q per stage, 16 stages from < 3q stay far below 2^48
    __device__ __forceinline__ V fwd_fix(V v) const { return v; }
    static __host__ __device__ constexpr int approx_stages(int) { return 0; }
    template <bool>
    __device__ __forceinline__ void fwd_bfly(V &X, V &Y, const TW W) const
    {
        const double T = mulmod(Y, W);
        Y = __dadd_rn(X, -T);
        X = __dadd_rn(X, T);
    }
    __device__ __forceinline__ u64 fwd_final(V v) const { return to_canonical(reduce(v)); }
    // inverse: the sum path doubles per stage; one reduction at the start of every full pass
    __device__ __forceinline__ V inv_fix(V v) const { return reduce(v); }
    template <int GBIT>
    __device__ __forceinline__ void inv_bfly(V &X, V &Y, const TW W) const
    {
        const double D = __dadd_rn(X, -Y);
        X = __dadd_rn(X, Y);
        Y = mulmod(D, W);
    }
    template <int GBIT>
    __device__ __forceinline__ void inv_bfly_last(V &X, V &Y) const
    {
        const double D = __dadd_rn(X, -Y);
        X = mulmod(__dadd_rn(X, Y), f.ninv);
        Y = mulmod(D, f.wlast);
    }
    __device__ __forceinline__ u64 inv_final(V v) const { return to_canonical(v); }
};

// compile-time loop: the body receives std::integral_constant<int, i>
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

// CNT consecutive twiddles starting at an index that is a multiple of CNT (a thread's butterflies of one stage use the
// twiddles CNT*a .. CNT*a + CNT-1): doubles come as 128-bit loads (two per instruction) instead of one LDG.64 each --
// in the last passes, where every lane pair reads its own twiddles, that halves the load instructions and the number
// of partially used sectors
template <int CNT>
__device__ __forceinline__ void load_tw(const double *__restrict__ p, double (&w)[CNT])
{
    if constexpr (CNT == 1) {
        w[0] = __ldg(p);
    } else {
#pragma unroll
        for (int i = 0; i < CNT; i += 2) {
            const double2 v = __ldg(reinterpret_cast<const double2 *>(p + i));
            w[i] = v.x;
            w[i + 1] = v.y;
        }
    }
}
template <int CNT>
__device__ __forceinline__ void load_tw(const ulonglong2 *__restrict__ p, ulonglong2 (&w)[CNT])
{
#pragma unroll
    for (int i = 0; i < CNT; ++i) w[i] = __ldg(p + i);
}

// ------------------------------------------------------------------ forward butterflies
// S stages on register bits S-1..0 (descending).  gbase = N + (global index of x[0]).
template <int LOGE, int S, int PLO, int PHI, class A>
__device__ __forceinline__ void fwd_stages(typename A::V (&x)[1 << LOGE], u32 gbase, const typename A::TW *__restrict__ tw, const A &ar)
{
    static_for<0, S>([&](auto ssc) {
        constexpr int ss = decltype(ssc)::value;
        constexpr int s = S - 1 - ss;
#pragma unroll
        for (int kh = 0; kh < (1 << (LOGE - S)); ++kh) {
            const u32 g = (S == LOGE) ? gbase : gbase + ((u32)kh << PHI);
            typename A::TW W[1 << ss];
            load_tw<(1 << ss)>(tw + (g >> (PLO + s + 1)), W);  // g has zeros in the register field: the index is a multiple of 2^ss
#pragma unroll
            for (int hi = 0; hi < (1 << ss); ++hi) {
#pragma unroll
                for (int lo = 0; lo < (1 << s); ++lo) {
                    const int k = (kh << S) | (hi << (s + 1)) | lo;
                    ar.template fwd_bfly<(ss < A::approx_stages(S))>(x[k], x[k | (1 << s)], W[hi]);
                }
            }
        }
    });
}

// full radix-16 passes of the forward transform, field position descending
// Loader protocol: ld.raw(i) issues the global load(s) of coefficient i and returns them untouched
// (type Loader::Raw); ld.fix(raw) is the arithmetic that turns them into a value < 4q.  Keeping
// the two apart lets pass 0 issue the loads of register set it+1 before it computes set it
// (Loader::PIPE; used where the extra registers do not spill: the plain transforms).
template <int LOGL, int LOGE, int PASS, class A, class Load>
__device__ __forceinline__ void fwd_full_passes(Load &ld, const typename A::TW *__restrict__ tw, u32 goff, const A &ar,
                                                typename A::V *sm)
{
    typedef NttShape<LOGL, LOGE> Sh;
    constexpr int E = 1 << LOGE;
    if constexpr (PASS < Sh::NFULL) {
        constexpr int P = LOGL - LOGE * (PASS + 1);
        typedef PassMap<LOGL, LOGE, LOGE, P, LOGL> M;
        if constexpr (PASS == 0 && Load::PIPE) {
            typename Load::Raw raw[2][E];
            {
                const u32 b0 = M::base(threadIdx.x);
#pragma unroll
                for (int k = 0; k < E; ++k) raw[0][k] = ld.raw(b0 + M::off(k));
            }
#pragma unroll
            for (int it = 0; it < Sh::ITER; ++it) {
                const u32 b = M::base(threadIdx.x + it * Sh::THREADS);
                if (it + 1 < Sh::ITER) {  // software pipeline: next set's loads fly during this set's butterflies
                    const u32 bn = M::base(threadIdx.x + (it + 1) * Sh::THREADS);
#pragma unroll
                    for (int k = 0; k < E; ++k) raw[(it + 1) & 1][k] = ld.raw(bn + M::off(k));
                }
                typename A::V x[E];
#pragma unroll
                for (int k = 0; k < E; ++k) x[k] = ar.from_load(ld.fix(raw[it & 1][k], b + M::off(k)));
                fwd_stages<LOGE, LOGE, P, LOGL>(x, goff + b, tw, ar);
#pragma unroll
                for (int k = 0; k < E; ++k) sm[swz(b + M::off(k))] = x[k];
            }
        } else if constexpr (PASS == 0) {
#pragma unroll 1
            for (int it = 0; it < Sh::ITER; ++it) {
                const u32 b = M::base(threadIdx.x + it * Sh::THREADS);
                typename Load::Raw raw[E];
#pragma unroll
                for (int k = 0; k < E; ++k) raw[k] = ld.raw(b + M::off(k));
                typename A::V x[E];
                ld.template fix_set<E, A>(raw, x, [&](int k) { return b + M::off(k); }, ar);
                fwd_stages<LOGE, LOGE, P, LOGL>(x, goff + b, tw, ar);
#pragma unroll
                for (int k = 0; k < E; ++k) sm[swz(b + M::off(k))] = x[k];
            }
        } else {
#pragma unroll 1
            for (int it = 0; it < Sh::ITER; ++it) {
                typename A::V x[E];
                const u32 b = M::base(threadIdx.x + it * Sh::THREADS);
#pragma unroll
                for (int k = 0; k < E; ++k) x[k] = ar.fwd_fix(sm[swz(b + M::off(k))]);
                fwd_stages<LOGE, LOGE, P, LOGL>(x, goff + b, tw, ar);
                // in-place: a thread only overwrites the slots it read itself
#pragma unroll
                for (int k = 0; k < E; ++k) sm[swz(b + M::off(k))] = x[k];
            }
        }
        __syncthreads();
        fwd_full_passes<LOGL, LOGE, PASS + 1>(ld, tw, goff, ar, sm);
    }
}

// One limb polynomial (or one 2^LOGL block of it), forward.
//  load: loader object (raw / fix, see above) for coefficient i of this CTA's block
//  store(i, v) receives the canonical result for position i of the block
//  goff = N_total + block offset (global index of local coefficient 0)
//  fetch(i) -> the epilogue operands of position i (issued for the whole register set before the
//  last stages run, so their latency hides behind the butterflies instead of serialising behind
//  the stores); store(i, v, ops) receives the canonical result and those operands
template <int LOGL, int LOGE, class A, class Load, class Fetch, class Store>
__device__ __forceinline__ void ntt_fwd_cta(Load load, Fetch fetch, Store store, const typename A::TW *__restrict__ tw, u32 goff,
                                            const A &ar, u64 *smraw)
{
    typedef NttShape<LOGL, LOGE> Sh;
    constexpr int E = 1 << LOGE;
    typename A::V *sm = reinterpret_cast<typename A::V *>(smraw);
    fwd_full_passes<LOGL, LOGE, 0>(load, tw, goff, ar, sm);
    // last pass: REM stages on the low bits; the other register bits are the top index bits
    constexpr int S = Sh::REM;
    constexpr int PHI = (S == LOGE) ? LOGL : LOGL - (LOGE - S);
    typedef PassMap<LOGL, LOGE, S, 0, PHI> M;
    // The epilogue operands of register set it + 1 are fetched element by element right after the store that consumed the
    // operands of set it (same registers): their latency then hides behind the remaining stores, the next set's shared-memory
    // reads and its butterflies, instead of behind one stage of butterflies (ncu: a third of the mod-down transform's
    // last-pass samples were long_scoreboard).  Only the first set's fetch is exposed.
    decltype(fetch(0u)) ops[E];
    {
        const u32 b0 = M::base(threadIdx.x);
#pragma unroll
        for (int k = 0; k < E; ++k) ops[k] = fetch(b0 + M::off(k));
    }
#pragma unroll 1
    for (int it = 0; it < Sh::ITER; ++it) {
        typename A::V x[E];
        const u32 b = M::base(threadIdx.x + it * Sh::THREADS);
        const u32 bn = M::base(threadIdx.x + (it + 1) * Sh::THREADS);
        const bool more = it + 1 < Sh::ITER;
#pragma unroll
        for (int k = 0; k < E; ++k) x[k] = ar.fwd_fix(sm[swz(b + M::off(k))]);
        fwd_stages<LOGE, S, 0, PHI>(x, goff + b, tw, ar);
#pragma unroll
        for (int k = 0; k < E; ++k) {
            store(b + M::off(k), ar.fwd_final(x[k]), ops[k]);
            if (more) ops[k] = fetch(bn + M::off(k));
        }
    }
}

// ------------------------------------------------------------------ inverse butterflies
// S stages on register bits 0..S-1 (ascending); global stage number = PLO + s.
template <int LOGE, int S, int PLO, int PHI, int LAST_BIT, class A>
__device__ __forceinline__ void inv_stages(typename A::V (&x)[1 << LOGE], u32 gbase, const typename A::TW *__restrict__ tw, const A &ar)
{
    static_for<0, S>([&](auto sc) {
        constexpr int s = decltype(sc)::value;
#pragma unroll
        for (int kh = 0; kh < (1 << (LOGE - S)); ++kh) {
            const u32 g = (S == LOGE) ? gbase : gbase + ((u32)kh << PHI);
            constexpr int CNT = 1 << (S - 1 - s);
            typename A::TW Wv[CNT];
            if (PLO + s != LAST_BIT) load_tw<CNT>(tw + (g >> (PLO + s + 1)), Wv);  // the index is a multiple of CNT (zeros in the register field)
#pragma unroll
            for (int hi = 0; hi < CNT; ++hi) {
                if (PLO + s == LAST_BIT) {
#pragma unroll
                    for (int lo = 0; lo < (1 << s); ++lo) {
                        const int k = (kh << S) | (hi << (s + 1)) | lo;
                        if (s == 0) ar.template inv_bfly_last<PLO>(x[k], x[k | 1]);
                        if (s == 1) ar.template inv_bfly_last<PLO + 1>(x[k], x[k | 2]);
                        if (s == 2) ar.template inv_bfly_last<PLO + 2>(x[k], x[k | 4]);
                        if (s == 3) ar.template inv_bfly_last<PLO + 3>(x[k], x[k | 8]);
                    }
                } else {
                    const typename A::TW W = Wv[hi];
#pragma unroll
                    for (int lo = 0; lo < (1 << s); ++lo) {
                        const int k = (kh << S) | (hi << (s + 1)) | lo;
                        if (s == 0) ar.template inv_bfly<PLO>(x[k], x[k | 1], W);
                        if (s == 1) ar.template inv_bfly<PLO + 1>(x[k], x[k | 2], W);
                        if (s == 2) ar.template inv_bfly<PLO + 2>(x[k], x[k | 4], W);
                        if (s == 3) ar.template inv_bfly<PLO + 3>(x[k], x[k | 8], W);
                    }
                }
            }
        }
    });
}

// full radix-16 passes of the inverse transform, field position ascending
template <int LOGL, int LOGE, int LAST_BIT, int PASS, class A, class Store>
__device__ __forceinline__ void inv_full_passes(Store &store, const typename A::TW *__restrict__ tw, u32 goff, const A &ar,
                                                typename A::V *sm)
{
    typedef NttShape<LOGL, LOGE> Sh;
    constexpr int E = 1 << LOGE;
    if constexpr (PASS < Sh::NFULL) {
        constexpr int P = Sh::REM + LOGE * PASS;
        typedef PassMap<LOGL, LOGE, LOGE, P, LOGL> M;
#pragma unroll 1
        for (int it = 0; it < Sh::ITER; ++it) {
            typename A::V x[E];
            const u32 b = M::base(threadIdx.x + it * Sh::THREADS);
#pragma unroll
            for (int k = 0; k < E; ++k) x[k] = ar.inv_fix(sm[swz(b + M::off(k))]);
            inv_stages<LOGE, LOGE, P, LOGL, LAST_BIT>(x, goff + b, tw, ar);
            if constexpr (PASS == Sh::NFULL - 1) {
                // the store receives the whole register set: it can issue the loads of everything it combines the set with
                // (parked halves, epilogue operands) before the arithmetic of the first element
                u32 idx[E];
#pragma unroll
                for (int k = 0; k < E; ++k) idx[k] = b + M::off(k);
                if constexpr (LAST_BIT >= 0) {
                    u64 y[E];
#pragma unroll
                    for (int k = 0; k < E; ++k) y[k] = ar.inv_final(x[k]);
                    store(idx, y);
                } else {
                    store(idx, x);
                }
            } else {
#pragma unroll
                for (int k = 0; k < E; ++k) sm[swz(b + M::off(k))] = x[k];
            }
        }
        if constexpr (PASS != Sh::NFULL - 1) {
            __syncthreads();
            inv_full_passes<LOGL, LOGE, LAST_BIT, PASS + 1>(store, tw, goff, ar, sm);
        }
    }
}

// One limb polynomial (block), inverse.  LAST_BIT = index of the final stage of the whole
// transform (logN-1) if this CTA performs it (store(idx[E], v[E]) receives the canonical values of a register set and their positions), else -1
// (SPLIT, integer policies only: values leave in [0,2q) (BIG) or [0, 2q*2^LOGL) (!BIG);
// ntt_inv_final_kernel finishes).  load(i) must return canonical residues.
template <int LOGL, int LOGE, int LAST_BIT, class A, class Load, class Store>
__device__ __forceinline__ void ntt_inv_cta(Load load, Store store, const typename A::TW *__restrict__ tw, u32 goff, const A &ar,
                                            u64 *smraw)
{
    typedef NttShape<LOGL, LOGE> Sh;
    constexpr int E = 1 << LOGE;
    typename A::V *sm = reinterpret_cast<typename A::V *>(smraw);
    constexpr int S = Sh::REM;
    constexpr int PHI = (S == LOGE) ? LOGL : LOGL - (LOGE - S);
    typedef PassMap<LOGL, LOGE, S, 0, PHI> M;
    if constexpr (Load::PIPE) {
        typename Load::Raw raw[2][E];
        {
            const u32 b0 = M::base(threadIdx.x);
#pragma unroll
            for (int k = 0; k < E; ++k) raw[0][k] = load.raw(b0 + M::off(k));
        }
#pragma unroll
        for (int it = 0; it < Sh::ITER; ++it) {
            const u32 b = M::base(threadIdx.x + it * Sh::THREADS);
            if (it + 1 < Sh::ITER) {
                const u32 bn = M::base(threadIdx.x + (it + 1) * Sh::THREADS);
#pragma unroll
                for (int k = 0; k < E; ++k) raw[(it + 1) & 1][k] = load.raw(bn + M::off(k));
            }
            typename A::V x[E];
#pragma unroll
            for (int k = 0; k < E; ++k) x[k] = ar.from_load(load.fix(raw[it & 1][k], b + M::off(k)));
            inv_stages<LOGE, S, 0, PHI, LAST_BIT>(x, goff + b, tw, ar);
#pragma unroll
            for (int k = 0; k < E; ++k) sm[swz(b + M::off(k))] = x[k];
        }
    } else {
#pragma unroll 1
        for (int it = 0; it < Sh::ITER; ++it) {
            const u32 b = M::base(threadIdx.x + it * Sh::THREADS);
            typename Load::Raw raw[E];
#pragma unroll
            for (int k = 0; k < E; ++k) raw[k] = load.raw(b + M::off(k));
            typename A::V x[E];
            load.template fix_set<E, A>(raw, x, [&](int k) { return b + M::off(k); }, ar);
            inv_stages<LOGE, S, 0, PHI, LAST_BIT>(x, goff + b, tw, ar);
#pragma unroll
            for (int k = 0; k < E; ++k) sm[swz(b + M::off(k))] = x[k];
        }
    }
    __syncthreads();
    inv_full_passes<LOGL, LOGE, LAST_BIT, 0>(store, tw, goff, ar, sm);
}

}  // namespace hegpu
