// mac_kernels.cuh -- the non-NTT kernels: key inner product, element-wise ops, tensor product, BSGS
// inner sums (unfused and double-hoisted fused), accumulations, repacking.  Included by hegpu.cu only
// (the NTT kernels and their job resolvers live in kernels.cuh, shared with the ntt_inst_*.cu units).
#pragma once
#include <type_traits>
#include "kernels.cuh"

namespace hegpu {

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

// K7 step 3 for lazy giant steps: the key inner products of `ngroups` rotations of the SAME batch are
// summed in the extended basis right here, out[b][c][i] = init[b][c][i] + sum_g <digits_g, key_g>[c][i],
// instead of writing one accumulator per rotation and summing them in a second pass (saves
// 2 * ngroups accumulator-sized HBM passes).  Each group's sum is reduced before it is added, exactly as
// the two-kernel form did, so the result is bit-identical.  Not for hoisted plans.
template <int LT>
__global__ void __launch_bounds__(256) ks_inner_sum_kernel(const KsParams P, const u64 *__restrict__ init, u64 *__restrict__ out,
                                                           const ModConst *__restrict__ mods)
{
    const u32 L = LT, n = P.n;
    const u32 x = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 i = blockIdx.y;
    const u32 b0 = blockIdx.z * KS_INNER_BCHUNK;
    const u32 b1 = min(P.B, b0 + KS_INNER_BCHUNK);
    if (x >= n) return;
    const u32 ki = (i == L) ? P.K - 1 : i;
    const ModConst m = mods[ki];
    const size_t kstride = (size_t)P.K * n;
    for (u32 b = b0; b < b1; ++b) {
        const size_t o0 = (((size_t)b * 2 + 0) * (L + 1) + i) * n + x, o1 = o0 + (size_t)(L + 1) * n;
        u64 s0 = init ? init[o0] : 0, s1 = init ? init[o1] : 0;
        // The products of up to GR groups share one 128-bit sum and ONE Montgomery reduction per component: GR * L products
        // of canonical residues stay below q * 2^64 while GR * L < 16 (q < 2^60), so the reduced value is canonical after one
        // conditional subtraction -- the same residue as the sum of the separately reduced group sums (bit-identical), for a
        // third of the reductions at three groups (they were ~28 % of this kernel's multiplier-pipe work).
        constexpr u32 GR = 15 / LT;
        for (u32 g0 = 0; g0 < P.ngroups; g0 += GR) {
            u64 h0 = 0, l0 = 0, h1 = 0, l1 = 0;
            const u32 g1 = min(P.ngroups, g0 + GR);
            for (u32 g = g0; g < g1; ++g) {
                const CtView &v = P.in[g];
                const u32 *pm = P.perm[g];
                const u32 xp = pm ? __ldg(pm + x) : x;
                const u32 xs = i < L ? xp : x;
                if (P.u0) s0 = addmod(s0, P.u0[((((size_t)g * P.B + b) * 2) * (L + 1) + i) * n + xp], m.q);
                const u64 *key = P.key[g] + (size_t)ki * n + x;
                const size_t e = (size_t)g * P.B + b;
#pragma unroll
                for (int j = 0; j < LT; ++j) {
                    const u64 d = ((u32)j == i) ? v.p[b * v.sb + P.target_poly * v.sp + j * v.sl + xs]
                                                : P.ext[((e * L + j) * (L + 1) + i) * n + x];
                    mac128(h0, l0, d, __ldg(key + (size_t)(2 * j) * kstride));      // L1-resident across the batch chunk
                    mac128(h1, l1, d, __ldg(key + (size_t)(2 * j + 1) * kstride));
                }
            }
            s0 = addmod(s0, mont_reduce(h0, l0, m), m.q);
            s1 = addmod(s1, mont_reduce(h1, l1, m), m.q);
        }
        out[o0] = s0;
        out[o1] = s1;
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

// Ciphertext::is_transparent: flags[b] = 1 when some word of the polynomials 1..size-1 of ciphertext b is non-zero
// (a ciphertext whose polynomials beyond c0 are all zero decrypts without the secret key).
__global__ void __launch_bounds__(256) nonzero_tail_kernel(const CtView v, u32 B, u32 polys, u32 L, u32 n, u32 *__restrict__ flags)
{
    const size_t per_b = (size_t)(polys - 1) * L * n;
    const size_t total = (size_t)B * per_b;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const u32 b = (u32)(idx / per_b);
        u32 r = (u32)(idx % per_b);
        const u32 x = r % n;
        r /= n;
        const u32 l = r % L, p = 1 + r / L;
        const bool nz = v.p[b * v.sb + p * v.sp + l * v.sl + x] != 0;
        if (__any_sync(__activemask(), nz) && nz && flags[b] == 0) flags[b] = 1;  // benign race: every writer stores 1
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
//   b_k[c]   = sum_j pi_k(digit_j) * key_k[j][c]  (+ P * pi_k(c0) for c = 0, limbs < L: the word of the
//              pre-scaled copy c0p = P * c0 (scale_c0_kernel) joins the Montgomery reduction of the sum)
//   u_g[c]   = sum_k b_k[c] * diag[g*n1+k]                            all in the extended basis
// A CTA owns one extended limb i and a tile of DH_TX = 32 coefficients; it stages the key words
// (n1 x 2L), the diagonal words (n1*n2) and the gather indices of that tile in shared memory ONCE
// and sweeps a chunk of the batch, so keys and diagonals leave HBM once per chunk and the rotated
// ciphertexts b_k never exist in HBM.  Inside the CTA every WARP owns whole ciphertexts (lane =
// coefficient): it walks the n1 baby steps, builds b_k in registers (the lifted digits are gathered
// from HBM/L2 -- a Galois permutation maps an aligned 32-word tile to an aligned tile, so a gather
// is one fully used 256-byte line -- and prefetched two baby steps ahead) and feeds it straight
// into the 2*n2 independent 128-bit accumulators of the giant steps: no barrier and no
// shared-memory round trip in the main loop, one Montgomery reduction per output word.  Baby step 0 is
// peeled and whole triples of baby steps run as one straight block (the three operand sets rotate by name).
// The kernel is bound by integer-multiply issue and dependent-issue latency (ncu: fmaheavy pipe 61-84 %
// busy depending on the variant), not by HBM: keys, diagonals, digits and outputs cross HBM once.
// More than 4 giant steps run as several launches of at most 4 (registers hold the accumulators; b_k is
// rebuilt per launch).  grid = ((N / 32) * (L+1), 1, batch chunks), block = 32 * DH_KG.
// ---------------------------------------------------------------------------------------
constexpr int DH_TX = 32, DH_KG = 8, DH_BCH = 64;
struct DhInnerParams {
    CtView in;             // input batch at level L
    const u64 *ext;        // [B][L][L+1][N] lifted digits of c1 (hoisted decomposition)
    const u64 *key[MAXB];  // Galois key of baby step k, Montgomery form; [0] unused
    const u32 *perm[MAXB]; // gather table of baby step k; [0] unused
    const u64 *diag;       // [n1*n2][Lcap][N] Montgomery form, limb L = special prime
    size_t diag_si;
    const u64 *c0p;        // [B][L][N]: P * c0 mod q_i (scale_c0_kernel), gathered into b_k[0] instead of a product per baby step
    u64 *u;                // [n2][B][2][L+1][N]
    u32 n1, n2, B, L, K, n;
    u32 g0, ng;            // this launch accumulates giant steps g0 .. g0+ng-1 (ng <= N2)
    u64 *bk;               // [B][L+1][n1][2][N] rotated ciphertexts b_k (lazy residues) of this chunk, or null: with more than N2
                           // giant steps the first launch writes them and the later launches read them back instead of gathering the
                           // digits and redoing the (n1-1) * 2L key products per launch
    u32 bk_mode;           // 0: not used, 1: compute and write, 2: read
    u32 stream_out;        // inner sums leave with evict-first stores (st.global.cs): they are > 1 GB per launch, are not read again
                           // by this kernel and would otherwise push the lifted digits that every CTA re-gathers out of L2
};
static inline size_t dh_inner_smem(u32 n1, u32 n2, u32 L)
{
    return ((size_t)n1 * 2 * L + (size_t)n1 * n2) * DH_TX * sizeof(u64) + (size_t)n1 * DH_TX * sizeof(u32);
}

// exact double of a 32-bit unsigned integer without the conversion unit: 2^52 + x has x in its low word
__device__ __forceinline__ double u32_to_f64(u32 x) { return __hiloint2double(0x43300000, (int)x) - 4503599627370496.0; }
// exact integer of a double in [0, 2^52)
__device__ __forceinline__ u64 f64_to_u64(double v) { return (u64)__double_as_longlong(v + 4503599627370496.0) & 0xFFFFFFFFFFFFFull; }

// Arithmetic policies of dh_inner_kernel (uniform per CTA = per extended limb).
//  DhArI64: any modulus.  Operands are 64-bit words, products accumulate in 128 bits (4 IMAD.WIDE.U32 +
//           three IADD3, see mac128), one Montgomery reduction per sum.
//  DhArF64: moduli below 2^40 (the 40-bit data primes).  Every word is split into two 20-bit limbs held
//           as doubles; the four limb products of a multiply-accumulate are < 2^40 and go into three
//           column sums (weights 1, 2^20, 2^40) with 4 DFMA on the otherwise idle FP64 pipe; up to 64
//           products stay below 2^47, far inside the 53-bit mantissa, so the sums are EXACT integers.
//           They are recombined in integers and Montgomery-reduced exactly like the I64 sums: both
//           policies return the same canonical residues, bit for bit.
struct DhArI64 {
    typedef u64 Opnd;
    struct Acc {
        u64 h, l;
    };
    static __device__ __forceinline__ u64 stage(u64 v) { return v; }
    static __device__ __forceinline__ Opnd from_word(u64 v) { return v; }
    static __device__ __forceinline__ Opnd from_staged(u64 v) { return v; }
    static __device__ __forceinline__ Acc zero() { return Acc{ 0, 0 }; }
    static __device__ __forceinline__ void mac(Acc &a, Opnd x, Opnd y) { mac128(a.h, a.l, x, y); }
    // b_k only feeds further products: a lazy value is enough (n1 <= 32 products of a value < 2.2 q < 2^62 and a
    // canonical diagonal word stay below 2^127)
    // `add` (canonical) joins the result: (hi:lo) + add * 2^64 reduces to (hi:lo) * 2^-64 + add.  With three key
    // products below q^2 <= q * 2^60 the result stays below 2.2 q.
    static __device__ __forceinline__ u64 reduce(const Acc &a, u64 add, const ModConst &m) { return mont_reduce_lazy(a.h + add, a.l, m); }
    static __device__ __forceinline__ u64 reduce_wide(const Acc &a, const ModConst &m)
    {
        // 32 products of a value below 2.2 q and a canonical word: sum < 4.4 * q * 2^64, reduced value < 5.4 q
        const u64 r = mont_reduce_lazy(a.h, a.l, m);
        return csub(csub(csub(r, 4 * m.q), 2 * m.q), m.q);
    }
};
struct DhArF64 {
    struct Opnd {
        double lo, hi;  // 20-bit limbs
    };
    struct Acc {
        double c0, c1, c2;
    };
    // staged words (keys, diagonals; canonical, below 2^40): the two 20-bit limbs as a pair of floats, exact, so
    // that a use costs one F2F.F64.F32 per limb on the conversion unit instead of mask / pair-move / DADD
    static __device__ __forceinline__ u64 stage(u64 v)
    {
        return (u64)__float_as_uint((float)((u32)v & 0xFFFFFu)) | ((u64)__float_as_uint((float)(u32)(v >> 20)) << 32);
    }
    static __device__ __forceinline__ Opnd from_word(u64 v) { return Opnd{ u32_to_f64((u32)v & 0xFFFFFu), u32_to_f64((u32)(v >> 20)) }; }
    static __device__ __forceinline__ Opnd from_staged(u64 p) { return Opnd{ (double)__uint_as_float((u32)p), (double)__uint_as_float((u32)(p >> 32)) }; }
    static __device__ __forceinline__ Acc zero() { return Acc{ 0.0, 0.0, 0.0 }; }
    static __device__ __forceinline__ void mac(Acc &a, const Opnd x, const Opnd y)
    {
        a.c0 = __fma_rn(x.lo, y.lo, a.c0);
        a.c1 = __fma_rn(x.lo, y.hi, a.c1);
        a.c1 = __fma_rn(x.hi, y.lo, a.c1);
        a.c2 = __fma_rn(x.hi, y.hi, a.c2);
    }
    static __device__ __forceinline__ u64 reduce(const Acc &a, u64 add, const ModConst &m)
    {
        const unsigned __int128 t = (unsigned __int128)f64_to_u64(a.c0) + ((unsigned __int128)f64_to_u64(a.c1) << 20) +
                                    ((unsigned __int128)f64_to_u64(a.c2) << 40);
        return mont_reduce_lazy((u64)(t >> 64) + add, (u64)t, m);  // t < 64 * 2^83 << q * 2^64; result < add + q + 1 < 2^41
    }
    static __device__ __forceinline__ u64 reduce_wide(const Acc &a, const ModConst &m) { return csub(reduce(a, 0, m), m.q); }
};

template <int LT, int N2, class Ar, int BK = 0>
__device__ __forceinline__ void dh_inner_body(const DhInnerParams &P, const ModConst &m, u32 i, u32 x0, u32 b0, u32 b1, u64 *skey, u64 *sdiag,
                                              u32 *sperm)
{
    constexpr u32 TX = DH_TX, NT = DH_TX * DH_KG, L = LT;
    const u32 n = P.n, n1 = P.n1;
    const u32 lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const u32 ki = (i == L) ? P.K - 1 : i;
    if constexpr (BK != 2) {
#pragma unroll 8
        for (u32 idx = threadIdx.x; idx < n1 * 2 * L * TX; idx += NT) {
            const u32 x = idx % TX, r = idx / TX, jc = r % (2 * L), k = r / (2 * L);
            skey[idx] = k ? Ar::stage(__ldg(P.key[k] + ((size_t)jc * P.K + ki) * n + x0 + x)) : 0;
        }
    }
#pragma unroll 8
    for (u32 idx = threadIdx.x; idx < n1 * N2 * TX; idx += NT) {
        const u32 x = idx % TX, r = idx / TX, g = r % N2, k = r / N2;
        sdiag[idx] = g < P.ng ? Ar::stage(__ldg(P.diag + (size_t)((P.g0 + g) * n1 + k) * P.diag_si + (size_t)i * n + x0 + x)) : 0;
    }
    if constexpr (BK != 2) {
        for (u32 idx = threadIdx.x; idx < n1 * TX; idx += NT) {
            const u32 x = idx % TX, k = idx / TX;
            sperm[idx] = k ? __ldg(P.perm[k] + x0 + x) : x0 + x;
        }
    }
    __syncthreads();
    const bool data_limb = i < L;
    const typename Ar::Opnd pm = Ar::from_word(m.pmont);  // 0 for the special prime: no c0 term, b_0[L] = 0
    for (u32 b = b0 + w; b < b1; b += DH_KG) {
        // operand streams of this ciphertext, fixed for all baby steps (only the gathered column
        // changes): src[j] = digit j lifted to limb i (digit i itself is c1's limb i), src[L] = c0's
        // limb i.  Baby step 0 uses the same words unpermuted: c0 = src[L], c1 = src[i].
        const u64 *src[LT + 1];
        {
            const u64 *cb = P.in.p + b * P.in.sb;
            const u64 *eb = P.ext + ((size_t)b * L * (L + 1) + i) * n;
#pragma unroll
            for (int j = 0; j < LT; ++j) src[j] = ((u32)j == i) ? cb + P.in.sp + j * P.in.sl : eb + (size_t)j * (L + 1) * n;
            src[LT] = P.c0p + ((size_t)b * L + (data_limb ? i : 0)) * n;  // P * c0, limb i
        }
        // b_k of this ciphertext and limb: [n1][2][N]
        u64 *bkp = BK ? P.bk + (((size_t)b * (L + 1) + i) * n1 * 2) * n + x0 + lane : nullptr;
        auto fetch = [&](u32 k, u64 (&d)[LT + 1]) {
            if constexpr (BK == 2) {
                d[0] = __ldcs(bkp + (size_t)(2 * k) * n);
                d[1] = __ldcs(bkp + (size_t)(2 * k + 1) * n);
            } else {
                const u32 xs = sperm[k * TX + lane];
#pragma unroll
                for (int j = 0; j < LT; ++j) d[j] = src[j][xs];
                d[LT] = data_limb ? src[LT][xs] : 0;
            }
        };
        typename Ar::Acc acc[N2][2];
#pragma unroll
        for (int g = 0; g < N2; ++g) acc[g][0] = acc[g][1] = Ar::zero();
        u64 d0[LT + 1], d1[LT + 1], d2[LT + 1];
        fetch(0, d0);
        if (n1 > 1) fetch(1, d1);
        // the tail of a baby step: b_k -> the 2 * n2 giant-step sums
        auto feed = [&](u32 k, const typename Ar::Opnd a0, const typename Ar::Opnd a1) {
            const u64 *dp = sdiag + (size_t)k * N2 * TX + lane;
#pragma unroll
            for (int g = 0; g < N2; ++g) {
                const typename Ar::Opnd dg = Ar::from_staged(dp[g * TX]);
                Ar::mac(acc[g][0], a0, dg);
                Ar::mac(acc[g][1], a1, dg);
            }
        };
        // baby step 0: b_0 = P * (c0, c1); component 0 is the gathered (pre-scaled) word itself
        {
            if (2 < n1) fetch(2, d2);
            if constexpr (BK == 2) {
                feed(0, Ar::from_word(d0[0]), Ar::from_word(d0[1]));
            } else {
                u64 w1 = d0[0];
#pragma unroll
                for (int j = 1; j < LT; ++j) w1 = ((u32)j == i) ? d0[j] : w1;
                typename Ar::Acc s1 = Ar::zero();
                Ar::mac(s1, Ar::from_word(w1), pm);
                const u64 r1 = Ar::reduce(s1, 0, m);
                if constexpr (BK == 1) {
                    __stcs(bkp, d0[LT]);
                    __stcs(bkp + n, r1);
                }
                feed(0, Ar::from_word(d0[LT]), Ar::from_word(r1));
            }
        }
        // baby step k >= 1: start the gathers of step k+2 into `nxt` (GUARD: only if it exists), then consume `cur`
        auto step = [&](auto guard, u32 k, const u64 (&cur)[LT + 1], u64 (&nxt)[LT + 1]) {
            if (!decltype(guard)::value || k + 2 < n1) fetch(k + 2, nxt);
            if constexpr (BK == 2) {
                feed(k, Ar::from_word(cur[0]), Ar::from_word(cur[1]));
            } else {
                typename Ar::Acc s0 = Ar::zero(), s1 = Ar::zero();
                const u64 *kp = skey + (size_t)k * 2 * L * TX + lane;
#pragma unroll
                for (int j = 0; j < LT; ++j) {
                    const typename Ar::Opnd dj = Ar::from_word(cur[j]);
                    Ar::mac(s0, dj, Ar::from_staged(kp[(2 * j) * TX]));
                    Ar::mac(s1, dj, Ar::from_staged(kp[(2 * j + 1) * TX]));
                }
                // b_k[0] = key products + P * pi_k(c0): the pre-scaled word joins the Montgomery reduction
                const u64 r0 = Ar::reduce(s0, cur[LT], m), r1 = Ar::reduce(s1, 0, m);
                if constexpr (BK == 1) {
                    __stcs(bkp + (size_t)(2 * k) * n, r0);
                    __stcs(bkp + (size_t)(2 * k + 1) * n, r1);
                }
                feed(k, Ar::from_word(r0), Ar::from_word(r1));
            }
        };
        // the three operand sets rotate by name, not by copying; whole triples run without range checks so that
        // the loop body is one straight block (no register moves where guarded paths would merge)
        u32 k = 1;
        for (; k + 4 < n1; k += 3) {
            step(std::false_type{}, k, d1, d0);
            step(std::false_type{}, k + 1, d2, d1);
            step(std::false_type{}, k + 2, d0, d2);
        }
        for (; k < n1; k += 3) {
            step(std::true_type{}, k, d1, d0);
            if (k + 1 < n1) step(std::true_type{}, k + 1, d2, d1);
            if (k + 2 < n1) step(std::true_type{}, k + 2, d0, d2);
        }
#pragma unroll
        for (int g = 0; g < N2; ++g) {
            if ((u32)g < P.ng) {
                u64 *up = P.u + ((((size_t)(P.g0 + g) * P.B + b) * 2) * (L + 1) + i) * n + x0 + lane;
                const u64 r0 = Ar::reduce_wide(acc[g][0], m), r1 = Ar::reduce_wide(acc[g][1], m);
                if (P.stream_out) {
                    __stcs(up, r0);
                    __stcs(up + (size_t)(L + 1) * n, r1);
                } else {
                    up[0] = r0;
                    up[(size_t)(L + 1) * n] = r1;
                }
            }
        }
    }
}

template <int LT, int N2, int BK>
__device__ __forceinline__ void dh_inner_cta(const DhInnerParams &P, const ModConst *__restrict__ mods, int use_f64, u64 *dh_smem)
{
    constexpr u32 TX = DH_TX, L = LT;
    // neighbouring CTAs take different limbs of the same tile: the CTAs that share an SM then mix the
    // integer-pipe policy (60-bit limbs) with the FP64-pipe policy (40-bit limbs)
    const u32 x0 = (blockIdx.x / (L + 1)) * TX;
    u32 i = blockIdx.x % (L + 1);
    if (const u32 period = (u32)use_f64 >> 8) {
        // L + 1 == 4 and `period` (the SM count) a multiple of 4: CTAs c and c + period land on the same SM in the first
        // wave and would take the same limb; rotate the limb by the wave number and order the limbs int, fp, int, fp
        const u32 j = (i + blockIdx.x / period) & 3u;
        i = j == 2 ? 3u : j == 3 ? 2u : j;
    }
    use_f64 &= 1;
    const u32 b0 = blockIdx.z * DH_BCH, b1 = min(P.B, b0 + DH_BCH);
    const ModConst m = mods[(i == L) ? P.K - 1 : i];
    u64 *skey = dh_smem;                                                // [n1][2L][TX]
    u64 *sdiag = skey + (size_t)P.n1 * 2 * L * TX;                       // [n1][N2][TX]
    u32 *sperm = reinterpret_cast<u32 *>(sdiag + (size_t)P.n1 * N2 * TX);  // [n1][TX]
    if (use_f64 && (m.q >> 40) == 0)
        dh_inner_body<LT, N2, DhArF64, BK>(P, m, i, x0, b0, b1, skey, sdiag, sperm);
    else
        dh_inner_body<LT, N2, DhArI64, BK>(P, m, i, x0, b0, b1, skey, sdiag, sperm);
}
template <int LT, int N2>
__global__ void __launch_bounds__(DH_TX * DH_KG, 2) dh_inner_kernel(const DhInnerParams P, const ModConst *__restrict__ mods, int use_f64)
{
    extern __shared__ __align__(16) u64 dh_smem[];
    dh_inner_cta<LT, N2, 0>(P, mods, use_f64, dh_smem);
}
// the launches of a matvec with more than N2 giant steps: BK = 1 (first launch) also writes the rotated ciphertexts b_k,
// BK = 2 (later launches) reads them back instead of gathering digits and redoing the key products.  Kernels of their own:
// as branches of dh_inner_kernel they cost the single-launch case 3 % (register allocation of the merged kernel).
template <int LT, int N2, int BK>
__global__ void __launch_bounds__(DH_TX * DH_KG, 2) dh_inner_bk_kernel(const DhInnerParams P, const ModConst *__restrict__ mods, int use_f64)
{
    extern __shared__ __align__(16) u64 dh_smem[];
    dh_inner_cta<LT, N2, BK>(P, mods, use_f64, dh_smem);
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
struct GatherU0 {  // optional: + sum_g pi_g(u0[g][b][0]) on component 0 (every limb)
    const u64 *u0;
    const u32 *perm[MAXG];
};
__global__ void __launch_bounds__(256) acc_group_sum_kernel(const u64 *__restrict__ acc, const u64 *__restrict__ init,
                                                            u64 *__restrict__ out, u32 groups, u32 B, u32 L, u32 K, u32 n,
                                                            const ModConst *__restrict__ mods, const GatherU0 G)
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
        if (G.u0 && r < (size_t)(L + 1) * n) {  // component 0
            const u32 x = (u32)(r % n);
            for (u32 g = 0; g < groups; ++g) s = addmod(s, G.u0[((size_t)g * B + b) * per_b + (r - x) + __ldg(G.perm[g] + x)], q);
        }
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

// P * c0 mod q_i for the fused double-hoisted inner kernel: out[b][i][x], canonical
__global__ void __launch_bounds__(256) scale_c0_kernel(const CtView in, u64 *__restrict__ out, u32 B, u32 L, u32 n,
                                                       const ModConst *__restrict__ mods)
{
    const size_t per_b = (size_t)L * n;
    const size_t total = (size_t)B * per_b;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const u32 b = (u32)(idx / per_b);
        const u32 r = (u32)(idx % per_b);
        const u32 x = r % n, i = r / n;
        const ModConst m = mods[i];
        u64 h = 0, lo = 0;
        mac128(h, lo, in.p[b * in.sb + i * in.sl + x], m.pmont);
        out[idx] = mont_reduce(h, lo, m);
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
