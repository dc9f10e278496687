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
//  * Shoup multiplication keeps butterflies at 10 IMAD each; lazy ranges are tracked at
//    compile time so primes below 2^58 (fwd) / 2^46 (inv) need no per-stage corrections;
//  * N = 32768 runs as two CTAs that each redo the first (stride N/2) stage from global
//    memory and then own one half (forward), or finish with one element-wise stage kernel
//    (inverse) -- see ntt_fwd_kernel / ntt_inv_kernel SPLIT.
#pragma once
#include "modarith.cuh"

namespace hegpu {

struct NttTables {
    const ulonglong2 *fwd;  // [K][N]  {psi^brev(i), shoup}   index m+i as in SURVEY 9.2
    const ulonglong2 *inv;  // [K][N]  {psi^-brev(i), shoup}
    const ModConst *mods;   // [K]
    const ulonglong2 *inv_last;  // [K] {inv[1]*N^-1, shoup}: bottom twiddle of the last INTT stage
    u32 n;                  // ring degree N
    u32 logn;
};

template <int LOGL>
struct NttShape {
    static constexpr int LSIZE = 1 << LOGL;
    static constexpr int SETS = LSIZE / 16;                       // 16-coefficient register sets per pass
    static constexpr int THREADS = SETS > 512 ? 512 : SETS;       // 512 threads leave 128 registers each
    static constexpr int ITER = SETS / THREADS;
    static constexpr int REM = (LOGL % 4 == 0) ? 4 : (LOGL % 4);  // stages of the partial pass
    static constexpr int NFULL = (LOGL - REM) / 4;                // number of full radix-16 passes
    static constexpr size_t SMEM = sizeof(u64) << LOGL;
};

__device__ __forceinline__ u32 swz(u32 idx) { return idx ^ ((idx >> 4) & 15u); }

// Position of the 16 coefficients a thread owns in a pass.  The low S bits of the register
// index k sit at index bits [PLO+S-1:PLO]; the remaining 4-S bits sit at [PHI+3-S:PHI];
// the thread id fills every other bit, low to high.
template <int LOGL, int S, int PLO, int PHI>
struct PassMap {
    __device__ static __forceinline__ u32 base(u32 t)
    {
        u32 lo = t & ((1u << PLO) - 1u);
        u32 mid = (t >> PLO) & ((1u << (PHI - PLO - S)) - 1u);
        u32 hi = (S == 4) ? 0u : (t >> (PHI - S));
        u32 r = lo | (mid << (PLO + S));
        if (S != 4) r |= hi << (PHI + 4 - S);
        return r;
    }
    __device__ static __forceinline__ u32 off(int k)
    {
        u32 r = (u32)(k & ((1 << S) - 1)) << PLO;
        if (S != 4) r |= (u32)(k >> S) << PHI;
        return r;
    }
};

// ------------------------------------------------------------------ forward butterflies
// S stages on register bits S-1..0 (descending).  gbase = N + (global index of x[0]).
template <int S, int PLO, int PHI>
__device__ __forceinline__ void fwd_stages(u64 (&x)[16], u32 gbase, const ulonglong2 *__restrict__ tw, u64 q)
{
    const u64 q2 = q << 1;
#pragma unroll
    for (int ss = 0; ss < S; ++ss) {
        const int s = S - 1 - ss;
#pragma unroll
        for (int kh = 0; kh < (1 << (4 - S)); ++kh) {
            const u32 g = (S == 4) ? gbase : gbase + ((u32)kh << PHI);
#pragma unroll
            for (int hi = 0; hi < (1 << ss); ++hi) {
                const ulonglong2 W = __ldg(tw + (g >> (PLO + s + 1)) + hi);
#pragma unroll
                for (int lo = 0; lo < (1 << s); ++lo) {
                    const int k = (kh << S) | (hi << (s + 1)) | lo;
                    const int k2 = k | (1 << s);
                    u64 T = mul_shoup_lazy(x[k2], W.x, W.y, q);
                    x[k2] = x[k] + q2 - T;
                    x[k] = x[k] + T;
                }
            }
        }
    }
}

// full radix-16 passes of the forward transform, field position descending
template <int LOGL, bool BIG, int PASS, class Load>
__device__ __forceinline__ void fwd_full_passes(Load &load, const ulonglong2 *__restrict__ tw, u32 goff, u64 q, u64 *sm)
{
    typedef NttShape<LOGL> Sh;
    if constexpr (PASS < Sh::NFULL) {
        constexpr int P = LOGL - 4 * (PASS + 1);
        typedef PassMap<LOGL, 4, P, LOGL> M;
#pragma unroll 1
        for (int it = 0; it < Sh::ITER; ++it) {
            u64 x[16];
            const u32 b = M::base(threadIdx.x + it * Sh::THREADS);
            if constexpr (PASS == 0) {
#pragma unroll
                for (int k = 0; k < 16; ++k) x[k] = load(b + M::off(k));
            } else {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    u64 v = sm[swz(b + M::off(k))];
                    x[k] = BIG ? csub(v, q << 3) : v;
                }
            }
            fwd_stages<4, P, LOGL>(x, goff + b, tw, q);
            // in-place: a thread only overwrites the slots it read itself
#pragma unroll
            for (int k = 0; k < 16; ++k) sm[swz(b + M::off(k))] = x[k];
        }
        __syncthreads();
        fwd_full_passes<LOGL, BIG, PASS + 1>(load, tw, goff, q, sm);
    }
}

// One limb polynomial (or one 2^LOGL block of it), forward.
//  load(i)  -> coefficient i of this CTA's block, any value < 3q
//  store(i, v) receives the canonical result for position i of the block
//  goff = N_total + block offset (global index of local coefficient 0)
// Lazy ranges (BIG, q < 2^60): values are < 8q at the start of a pass and grow by 2q per
// stage (< 16q < 2^64); !BIG (q < 2^58): no corrections, < 33q at the end.
template <int LOGL, bool BIG, class Load, class Store>
__device__ __forceinline__ void ntt_fwd_cta(Load load, Store store, const ulonglong2 *__restrict__ tw, u32 goff,
                                            const ModConst &m, u64 *sm)
{
    typedef NttShape<LOGL> Sh;
    const u64 q = m.q;
    fwd_full_passes<LOGL, BIG, 0>(load, tw, goff, q, sm);
    // last pass: REM stages on the low bits; the other register bits are the top index bits
    constexpr int S = Sh::REM;
    constexpr int PHI = (S == 4) ? LOGL : LOGL - (4 - S);
    typedef PassMap<LOGL, S, 0, PHI> M;
#pragma unroll 1
    for (int it = 0; it < Sh::ITER; ++it) {
        u64 x[16];
        const u32 b = M::base(threadIdx.x + it * Sh::THREADS);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            u64 v = sm[swz(b + M::off(k))];
            x[k] = BIG ? csub(v, q << 3) : v;
        }
        fwd_stages<S, 0, PHI>(x, goff + b, tw, q);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            u64 v = x[k];
            if (BIG) {
                v = csub(v, q << 3);
                v = csub(v, q << 2);
                v = csub(v, q << 1);
                v = csub(v, q);
            } else {
                v = barrett64(v, m);
            }
            store(b + M::off(k), v);
        }
    }
}

// ------------------------------------------------------------------ inverse butterflies
// S stages on register bits 0..S-1 (ascending).  BIT0 = global bit position of register
// bit 0 (PLO); for !BIG the bound of every value entering global stage c is 2q * 2^c.
template <int S, int PLO, int PHI, bool BIG, int LAST_BIT>
__device__ __forceinline__ void inv_stages(u64 (&x)[16], u32 gbase, const ulonglong2 *__restrict__ tw, const ModConst &m,
                                           const ulonglong2 wlast)
{
    const u64 q = m.q, q2 = q << 1;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int gbit = PLO + s;  // global stage number
#pragma unroll
        for (int kh = 0; kh < (1 << (4 - S)); ++kh) {
            const u32 g = (S == 4) ? gbase : gbase + ((u32)kh << PHI);
#pragma unroll
            for (int hi = 0; hi < (1 << (S - 1 - s)); ++hi) {
                ulonglong2 W = __ldg(tw + (g >> (PLO + s + 1)) + hi);
                if (gbit == LAST_BIT) W = wlast;
#pragma unroll
                for (int lo = 0; lo < (1 << s); ++lo) {
                    const int k = (kh << S) | (hi << (s + 1)) | lo;
                    const int k2 = k | (1 << s);
                    const u64 X = x[k], Y = x[k2];
                    u64 Sm, D;
                    if (BIG) {
                        Sm = csub(X + Y, q2);
                        D = X + q2 - Y;
                    } else {
                        Sm = X + Y;
                        D = X + (q << (gbit + 1)) - Y;
                    }
                    if (gbit == LAST_BIT) {
                        x[k] = mul_shoup(Sm, m.ninv, m.ninv_sh, q);
                        x[k2] = mul_shoup(D, W.x, W.y, q);
                    } else {
                        x[k] = Sm;
                        x[k2] = mul_shoup_lazy(D, W.x, W.y, q);
                    }
                }
            }
        }
    }
}

// full radix-16 passes of the inverse transform, field position ascending
template <int LOGL, bool BIG, int LAST_BIT, int PASS, class Store>
__device__ __forceinline__ void inv_full_passes(Store &store, const ulonglong2 *__restrict__ tw, u32 goff, const ModConst &m,
                                                const ulonglong2 wlast, u64 *sm)
{
    typedef NttShape<LOGL> Sh;
    if constexpr (PASS < Sh::NFULL) {
        constexpr int P = Sh::REM + 4 * PASS;
        typedef PassMap<LOGL, 4, P, LOGL> M;
#pragma unroll 1
        for (int it = 0; it < Sh::ITER; ++it) {
            u64 x[16];
            const u32 b = M::base(threadIdx.x + it * Sh::THREADS);
#pragma unroll
            for (int k = 0; k < 16; ++k) x[k] = sm[swz(b + M::off(k))];
            inv_stages<4, P, LOGL, BIG, LAST_BIT>(x, goff + b, tw, m, wlast);
            if constexpr (PASS == Sh::NFULL - 1) {
#pragma unroll
                for (int k = 0; k < 16; ++k) store(b + M::off(k), x[k]);
            } else {
#pragma unroll
                for (int k = 0; k < 16; ++k) sm[swz(b + M::off(k))] = x[k];
            }
        }
        if constexpr (PASS != Sh::NFULL - 1) {
            __syncthreads();
            inv_full_passes<LOGL, BIG, LAST_BIT, PASS + 1>(store, tw, goff, m, wlast, sm);
        }
    }
}

// One limb polynomial (block), inverse.  LAST_BIT = index of the final stage of the whole
// transform (logN-1) if this CTA performs it (store receives canonical values), else -1
// (SPLIT: values leave in [0,2q) (BIG) or [0, 2q*2^LOGL) (!BIG); ntt_inv_final_kernel finishes).
// load(i) must return canonical residues.
template <int LOGL, bool BIG, int LAST_BIT, class Load, class Store>
__device__ __forceinline__ void ntt_inv_cta(Load load, Store store, const ulonglong2 *__restrict__ tw, u32 goff,
                                            const ModConst &m, const ulonglong2 wlast, u64 *sm)
{
    typedef NttShape<LOGL> Sh;
    constexpr int S = Sh::REM;
    constexpr int PHI = (S == 4) ? LOGL : LOGL - (4 - S);
    typedef PassMap<LOGL, S, 0, PHI> M;
#pragma unroll 1
    for (int it = 0; it < Sh::ITER; ++it) {
        u64 x[16];
        const u32 b = M::base(threadIdx.x + it * Sh::THREADS);
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = load(b + M::off(k));
        inv_stages<S, 0, PHI, BIG, LAST_BIT>(x, goff + b, tw, m, wlast);
#pragma unroll
        for (int k = 0; k < 16; ++k) sm[swz(b + M::off(k))] = x[k];
    }
    __syncthreads();
    inv_full_passes<LOGL, BIG, LAST_BIT, 0>(store, tw, goff, m, wlast, sm);
}

}  // namespace hegpu
