// imma_kernels.cuh -- OPT-IN, experimental: the fused double-hoisted inner sums (dh_inner_kernel, mac_kernels.cuh) as
// exact integer matrix products on the warp-level integer MMA units (mma.sync.m16n8k32 u8 x u8 -> s32).
//
// The north star rules tensor cores out for this path ("modular 64-bit arithmetic is not a dense FP contraction"), so
// this is NOT the default and the headline is measured without it.  It exists because the measurement says the inner
// sums are bound by 32x32-bit multiplier issue (IMAD.WIDE: 32 per clock per SM), not by HBM, and VERDICT r1 asked for
// a go/no-go on the 8-bit-limb form: for one extended limb i and one coefficient x
//      u_g[c][b] = sum_k ( sum_j pi_k(digit_j)[b] * key_k[j][c]  +  [c = 0] P * pi_k(c0)[b] ) * diag[g n1 + k]      (mod q_i)
//                = sum_e D[b][e] * W[e][(g, c)]   with  W = key (.) diag  (resp. diag, P * diag) pre-multiplied mod q_i,
// a [batch x n1*(L+1)] x [n1*(L+1) x 2*n2] integer product per (i, x): M = ciphertexts, N = 8 = 4 giant steps x 2
// components, K = n1 * (L+1).  Words are split into 8-bit limbs; the limb products of equal weight share one s32
// accumulator (column sums stay below 128 * 8 * 255^2 < 2^27), the 15 (9) column sums are recombined to the exact
// 128-bit integer and reduced with one Barrett step: the result is the same canonical residue dh_inner_kernel writes,
// bit for bit.  W is built once per (diagonals, keys) pair, already in the B-fragment order of the instruction.
#pragma once
#include "kernels.cuh"

namespace hegpu {

constexpr int IM_MT = 16;                    // ciphertexts per MMA tile (M)
constexpr int IM_WL = 8;                     // limb slots per k-step in the W layout

struct ImmaParams {
    CtView in;             // input batch at level L
    const u64 *ext;        // [B][L][L+1][N] lifted digits of c1
    const u64 *c0p;        // [B][L][N] P * c0 mod q_i
    const u32 *perm[MAXB]; // gather table of baby step k; [0] unused (identity)
    const u32 *w;          // [L+1][N][ksteps][IM_WL][32][2] B fragments of W (bytes), built by dh_imma_prep_kernel
    u64 *u;                // [n2][B][2][L+1][N]
    u32 n1, B, L, K, n;
    u32 g0, ng;            // giant steps g0 .. g0+ng-1 of this launch (ng <= 4)
    u32 ksteps;            // ceil(n1 * (L+1) / 32)
};

struct ImmaPrepParams {
    const u64 *key[MAXB];  // Galois key of baby step k, Montgomery form; [0] unused
    const u64 *diag;       // [n1*n2][Lcap][N] plain residues, limb L = special prime
    size_t diag_si;
    u32 *w;
    u32 n1, L, K, n, g0, ng, ksteps;
};

// one thread per (limb i, coefficient x, entry e, column): W word -> 8 bytes scattered into the fragment layout
__global__ void __launch_bounds__(256) dh_imma_prep_kernel(const ImmaPrepParams P, const ModConst *__restrict__ mods)
{
    const u32 L = P.L, E = P.ksteps * 32;
    const size_t total = (size_t)(L + 1) * P.n * E * 8;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const u32 col = (u32)(idx & 7), e = (u32)((idx >> 3) % E);
        const size_t ix = (idx >> 3) / E;
        const u32 x = (u32)(ix % P.n), i = (u32)(ix / P.n);
        const u32 g = col >> 1, cc = col & 1, k = e / (L + 1), j = e % (L + 1);
        const ModConst m = mods[i == L ? P.K - 1 : i];
        u64 wv = 0;
        if (g < P.ng && k < P.n1) {
            const u64 d = P.diag[(size_t)((P.g0 + g) * P.n1 + k) * P.diag_si + (size_t)i * P.n + x];  // canonical
            if (j == L) {  // the P * pi_k(c0) term: component 0 only, data limbs only
                if (cc == 0 && i < L) wv = d;
            } else if (k == 0) {  // b_0 = P * (c0, c1): c1's own limb i carries P * diag
                if (j == i && cc == 1 && i < L) wv = mont_reduce(__umul64hi(m.pmont, d), m.pmont * d, m);
            } else {
                const u64 kw = P.key[k][((size_t)(2 * j + cc) * P.K + (i == L ? P.K - 1 : i)) * P.n + x];  // key * 2^64 mod q
                wv = mont_reduce(__umul64hi(kw, d), kw * d, m);
            }
        }
        const u32 ks = e >> 5, el = e & 31, q = el >> 4, tg = (el & 15) >> 2, t = el & 3, lane = (col << 2) | tg;
        unsigned char *base = reinterpret_cast<unsigned char *>(P.w) + (((ix * P.ksteps + ks) * IM_WL) * 32 + lane) * 8 + q * 4 + t;
#pragma unroll
        for (int b = 0; b < IM_WL; ++b) base[(size_t)b * 32 * 8] = (unsigned char)(wv >> (8 * b));
    }
}

__device__ __forceinline__ u32 prmt(u32 a, u32 b, u32 sel)
{
    u32 d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// byte a of four 32-bit words -> one register per byte position: out[a] = {w0.a, w1.a, w2.a, w3.a}
__device__ __forceinline__ void transpose4x4(u32 w0, u32 w1, u32 w2, u32 w3, u32 (&out)[4])
{
    const u32 t0 = prmt(w0, w1, 0x5140), t1 = prmt(w2, w3, 0x5140), t2 = prmt(w0, w1, 0x7362), t3 = prmt(w2, w3, 0x7362);
    out[0] = prmt(t0, t1, 0x5410);
    out[1] = prmt(t0, t1, 0x7632);
    out[2] = prmt(t2, t3, 0x5410);
    out[3] = prmt(t2, t3, 0x7632);
}
__device__ __forceinline__ void imma(int (&c)[4], const u32 (&a)[4], u32 b0, u32 b1)
{
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// shared memory: two stages of {16 ciphertexts x 32 entries x 16 coefficients of D (64 KiB), this k-step's W fragments of
// the 16 coefficients (32 KiB)} + the gather tables: one CTA of 16 warps per SM
static inline size_t dh_imma_smem(u32 n1, u32 tx)
{
    return 2 * ((size_t)IM_MT * 32 * tx * sizeof(u64) + (size_t)tx * IM_WL * 64 * sizeof(u32)) + (size_t)n1 * tx * sizeof(u32);
}

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, bool valid)
{
    const u32 s = (u32)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(valid ? 8 : 0));  // src-size 0: zero fill
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const u32 s = (u32)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
// staged word (ciphertext ct, entry e, coefficient xx).  64-bit shared-memory accesses are served per HALF-warp against
// 16 banks of 8 bytes, so each half-warp of the staging stores (TX = 16: 16 coefficients of one entry; TX = 8: 8
// coefficients x 2 entries) and of the fragment loads (fixed coefficient, lanes = 4 ciphertexts x 4 entry groups) must
// cover all 16 banks: the coefficient index is XORed with ciphertext and entry-group bits and, for TX = 8 (rows of half a
// bank span), the entry order is permuted so that the row parity follows entry bit 3.
template <int TX>
__device__ __forceinline__ u32 im_word(u32 ct, u32 e, u32 xx)
{
    if (TX == 16) return ct * 512 + e * 16 + (xx ^ ((ct & 3u) | (((e >> 2) & 3u) << 2)));
    const u32 row = ((e >> 3) & 1u) | ((e & 3u) << 1) | (((e >> 2) & 1u) << 3) | (((e >> 4) & 1u) << 4);
    return ct * 256 + row * 8 + (xx ^ ((ct & 3u) | (((e >> 2) & 1u) << 2)));
}

// grid = ((N / TX) * (L+1), 1, 1), block = 32 * TX: warp w owns coefficient x0 + w of limb i for the whole batch
// (TX = 16: one CTA of 16 warps per SM, 128-byte gather rows; TX = 8: two CTAs of 8 warps, 64-byte rows).
// Steps (ciphertext tile, k-step) are software-pipelined: the asynchronous copies of step s+1 (gathered words and W
// fragments, global -> shared without registers) run while the MMAs of step s issue.
template <int LT, int LIMBS, int TX>
__device__ __forceinline__ void dh_imma_body(const ImmaParams &P, const ModConst &m, u32 i, u32 x0, u64 *dbuf, u32 *wbuf, u32 *sperm)
{
    constexpr u32 L = LT, NS = 2 * LIMBS - 1, IM_TX = TX, IM_THREADS = TX * 32, IM_DWORDS = IM_MT * 32 * TX, IM_WWORDS = TX * IM_WL * 64;
    const u32 n = P.n, tid = threadIdx.x, lane = tid & 31u, w = tid >> 5, g8 = lane >> 2, tg = lane & 3u;
    for (u32 q = tid; q < P.n1 * IM_TX; q += IM_THREADS) {
        const u32 k = q / IM_TX, xx = q % IM_TX;
        sperm[q] = k ? __ldg(P.perm[k] + x0 + xx) : x0 + xx;
    }
    // the staging role of this thread: entry se of every k-step, coefficient sxx, all 16 ciphertexts of the tile
    // TX = 8: lane bit 3 -> entry bit 3 and lane bit 4 -> entry bit 2, so that a half-warp stores two entries of different row parity
    const u32 sxx = tid & (IM_TX - 1),
              se = TX == 16 ? tid / 16 : (((lane >> 3) & 1u) << 3) | (((lane >> 4) & 1u) << 2) | (w & 3u) | ((w >> 2) << 4), sj = se % (L + 1);
    const u64 *sbase;
    size_t sstride;
    bool svalid = true;
    if (sj == L) {  // P * c0
        sbase = P.c0p + (size_t)(i < L ? i : 0) * n;
        sstride = (size_t)L * n;
        svalid = i < L;
    } else if (sj == i) {  // digit i at limb i is c1's own limb
        sbase = P.in.p + P.in.sp + (size_t)sj * P.in.sl;
        sstride = P.in.sb;
    } else {
        sbase = P.ext + ((size_t)sj * (L + 1) + i) * n;
        sstride = (size_t)L * (L + 1) * n;
    }
    const u32 *wsrc = P.w + ((size_t)i * n + x0) * P.ksteps * IM_WL * 64;
    const u32 tiles = (P.B + IM_MT - 1) / IM_MT, steps = tiles * P.ksteps;
    __syncthreads();  // sperm
    // shared-memory destinations of this thread's copies, as 32-bit shared addresses relative to the stage base: the XOR
    // term of im_word depends on ct & 3 only, so four bases + compile-time ciphertext strides cover the 16 ciphertexts
    const u32 dsh = (u32)__cvta_generic_to_shared(dbuf), wsh = (u32)__cvta_generic_to_shared(wbuf);
    u32 dofs[4];
#pragma unroll
    for (u32 cc = 0; cc < 4; ++cc) dofs[cc] = dsh + im_word<TX>(cc, se, sxx) * 8;
    u32 wofs[IM_WWORDS / 4 / IM_THREADS];
    size_t wsofs[IM_WWORDS / 4 / IM_THREADS];
#pragma unroll
    for (u32 it = 0; it < IM_WWORDS / 4 / IM_THREADS; ++it) {
        const u32 idx = tid + IM_THREADS * it, ww = idx / (IM_WL * 16), o = idx % (IM_WL * 16);
        wofs[it] = wsh + (ww * IM_WL * 64 + o * 4) * 4;
        wsofs[it] = (size_t)ww * P.ksteps * IM_WL * 64 + o * 4;
    }
    auto issue = [&](u32 step, u32 ks, u32 b0) {
        const u32 st = step & 1u;
        const u32 k = (ks * 32 + se) / (L + 1);
        const bool ok = svalid && k < P.n1;
        const u32 off = ok ? sperm[k * IM_TX + sxx] : 0;
        const u64 *p = sbase + (size_t)b0 * sstride + off;
        const u32 dst = st * (IM_DWORDS * 8);
        if (ok && b0 + IM_MT <= P.B) {  // the common case: every ciphertext of the tile exists, plain 8-byte copies
#pragma unroll
            for (u32 ct = 0; ct < IM_MT; ++ct) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dofs[ct & 3u] + dst + (ct >> 2) * (4 * 32 * IM_TX * 8)), "l"(p));
                p += sstride;
            }
        } else {
#pragma unroll
            for (u32 ct = 0; ct < IM_MT; ++ct) {
                const bool v = ok && b0 + ct < P.B;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dofs[ct & 3u] + dst + (ct >> 2) * (4 * 32 * IM_TX * 8)),
                             "l"(v ? p : sbase), "r"(v ? 8 : 0));  // src-size 0: zero fill
                p += sstride;
            }
        }
        const u32 *ws = wsrc + (size_t)ks * IM_WL * 64;
#pragma unroll
        for (u32 it = 0; it < IM_WWORDS / 4 / IM_THREADS; ++it)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(wofs[it] + st * (IM_WWORDS * 4)), "l"(ws + wsofs[it]));
        asm volatile("cp.async.commit_group;");
    };
    issue(0, 0, 0);
    int acc[NS][4];
    u32 ks = 0, b0 = 0;
    for (u32 step = 0; step < steps; ++step) {
        const u32 st = step & 1u;
        const u32 ksn = ks + 1 == P.ksteps ? 0 : ks + 1, b0n = ks + 1 == P.ksteps ? b0 + IM_MT : b0;  // the next step
        if (ks == 0) {
#pragma unroll
            for (int s = 0; s < (int)NS; ++s) acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = 0;
        }
        if (step + 1 < steps) {
            issue(step + 1, ksn, b0n);
            asm volatile("cp.async.wait_group 1;");
        } else {
            asm volatile("cp.async.wait_group 0;");
        }
        __syncthreads();
        // A fragments of coefficient w: rows g8 / g8+8 (ciphertexts), entries tg*4 .. +3 and +16
        const u64 *dcur = dbuf + (size_t)st * IM_DWORDS;
        u32 A[8][4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {  // h = 2*q + rowhalf: register index of the fragment
            const u32 ct = g8 + 8 * (h & 1), e0 = tg * 4 + 16 * (h >> 1);
            const u64 w0 = dcur[im_word<TX>(ct, e0, w)], w1 = dcur[im_word<TX>(ct, e0 + 1, w)], w2 = dcur[im_word<TX>(ct, e0 + 2, w)],
                      w3 = dcur[im_word<TX>(ct, e0 + 3, w)];
            u32 lo[4], hi[4];
            transpose4x4((u32)w0, (u32)w1, (u32)w2, (u32)w3, lo);
            A[0][h] = lo[0];
            A[1][h] = lo[1];
            A[2][h] = lo[2];
            A[3][h] = lo[3];
            if (LIMBS > 4) {
                transpose4x4((u32)(w0 >> 32), (u32)(w1 >> 32), (u32)(w2 >> 32), (u32)(w3 >> 32), hi);
                A[4][h] = hi[0];
                A[5][h] = hi[1];
                A[6][h] = hi[2];
                A[7][h] = hi[3];
            }
        }
        const uint2 *bf = reinterpret_cast<const uint2 *>(wbuf + (size_t)st * IM_WWORDS + (size_t)w * IM_WL * 64) + lane;
#pragma unroll
        for (int b = 0; b < LIMBS; ++b) {
            const uint2 bb = bf[b * 32];
#pragma unroll
            for (int a = 0; a < LIMBS; ++a) imma(acc[a + b], A[a], bb.x, bb.y);
        }
        __syncthreads();  // stage st may be overwritten by the copies of step + 2
        if (ks + 1 == P.ksteps) {
            // epilogue: column sums -> exact 128-bit integer -> canonical residue
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                // groups of four column sums fit 64 bits (each sum < 2^27): four shifted adds per group, then one
                // 128-bit shift-add per group instead of one per column
                unsigned __int128 V = 0;
#pragma unroll
                for (int gq = (int)(NS - 1) / 4; gq >= 0; --gq) {
                    u64 part = 0;
#pragma unroll
                    for (int s = 3; s >= 0; --s)
                        if (gq * 4 + s < (int)NS) part = (part << 8) + (u64)(u32)acc[gq * 4 + s][r];
                    V = (V << 32) + part;
                }
                const u32 b = b0 + g8 + 8 * (r >> 1), col = tg * 2 + (r & 1), g = col >> 1, cc = col & 1;
                if (b < P.B && g < P.ng)
                    P.u[((((size_t)(P.g0 + g) * P.B + b) * 2 + cc) * (L + 1) + i) * n + x0 + w] = barrett128((u64)(V >> 64), (u64)V, m);
            }
        }
        ks = ksn;
        b0 = b0n;
    }
}

template <int LT, int TX>
__global__ void __launch_bounds__(TX * 32, TX == 16 ? 1 : 2) dh_imma_kernel(const ImmaParams P, const ModConst *__restrict__ mods)
{
    extern __shared__ __align__(16) u64 im_smem[];
    constexpr u32 L = LT;
    const u32 x0 = (blockIdx.x / (L + 1)) * TX, i = blockIdx.x % (L + 1);
    const ModConst m = mods[(i == L) ? P.K - 1 : i];
    u64 *dbuf = im_smem;                                                   // [2][16 * 32 * TX]
    u32 *wbuf = reinterpret_cast<u32 *>(dbuf + 2 * IM_MT * 32 * TX);        // [2][TX * IM_WL * 64]
    u32 *sperm = wbuf + 2 * TX * IM_WL * 64;                                // [n1][TX]
    if ((m.q >> 40) == 0)
        dh_imma_body<LT, 5, TX>(P, m, i, x0, dbuf, wbuf, sperm);
    else
        dh_imma_body<LT, 8, TX>(P, m, i, x0, dbuf, wbuf, sperm);
}

}  // namespace hegpu
