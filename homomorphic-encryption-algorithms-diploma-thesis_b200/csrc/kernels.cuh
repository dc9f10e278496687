// kernels.cuh -- NTT job resolvers and NTT kernels of the CKKS evaluator (sm_100a); shared by hegpu.cu and
// the ntt_inst_*.cu translation units.  The multiply-accumulate / element-wise kernels are in mac_kernels.cuh.
//
//  K1/K2  ntt_fwd_kernel / ntt_inv_kernel (+ ntt_inv_final_kernel for N = 32768), park kernels (N = 16384)
//  K6     Galois permutation: fused as a gather into the key-switch INTT load, the key
//         inner product (digit i == j) and the mod-down epilogue -- never a separate pass
//  K7     key-switch: ks INTT -> lift+NTT -> key inner product -> INTT(+half) -> mod-down NTT
//  K8     rescale: INTT(+half) -> mod-down NTT (same two kernels as the end of K7)
//  (mac_kernels.cuh: K3-K5 element-wise / tensor, K7 step 3 ks_inner[_sum], K10 bsgs_inner / dh_inner,
//   K11 fixup)
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

// v mod m for a canonical residue v of the modulus `from`: nothing when from <= m, one conditional
// subtraction when from < 2m (both 60-bit primes of a chain), Barrett otherwise.  Uniform per CTA.
__device__ __forceinline__ u64 rebase(u64 v, u64 from, const ModConst &m)
{
    if (from <= m.q) return v;
    return (from >> 1) < m.q ? csub(v, m.q) : barrett64(v, m);
}

// Every resolver splits its work in two: resolve(j) runs ONCE per CTA and turns the job id into row pointers and the
// constants of the limb (type R; the kernels keep it in shared memory when it is large), the per-coefficient functions
// only index those rows.  In the first form every per-coefficient call redid the index arithmetic (integer divisions,
// 64-bit stride products, key / constant table look-ups): ptxas did not hoist it out of the unrolled passes under the
// 80-register budget -- the loader and epilogue passes of the mod-down transform were 1300-1600 instructions per
// iteration next to 340 for a plain radix-8 pass, with 9-14 division sequences each, and stall_no_instruction was its
// second largest stall (160-190 KB of code per kernel).
//   u32  mod(j)                         modulus index of job j
//   R    resolve(j)                     rows + constants of job j
//   u64  load_raw(r, i, m)              memory only (input coefficient i)
//   u64  load_fix(r, v, m)              arithmetic only (-> value below q, or below 16 q for lazy inputs)
//   inverse jobs:  void store(r, i, x, m)
//   forward jobs:  Ops fetch(r, i, m);  void store(r, i, x, m, ops)     (epilogue operands fetched before the last stages)

// the same for E values at once: the branch on the pair of moduli is taken once, not once per coefficient (ptxas keeps
// the per-coefficient form as a branch tree in front of every load's arithmetic)
template <int E>
__device__ __forceinline__ void rebase_set(u64 (&v)[E], u64 from, const ModConst &m)
{
    if (from <= m.q) return;
    if ((from >> 1) < m.q) {
#pragma unroll
        for (int k = 0; k < E; ++k) v[k] = csub(v[k], m.q);
    } else {
#pragma unroll
        for (int k = 0; k < E; ++k) v[k] = barrett64(v[k], m);
    }
}
// canonical residues of the modulus `from` as doubles congruent mod q (FP64 butterfly policy, q < 2^43), magnitude below
// q/2 + 2^30 or below 2q: a word below 2^51 converts exactly (and is reduced once if `from` is far above q); a 60-bit word is
// split at bit 30 and hi * 2^30 is reduced with the exact FP64 product -- 2 conversions + 7 FP64 operations instead of the
// 64-bit Barrett reduction (4 wide multiplies + ~15 integer instructions) followed by a conversion
template <int E>
__device__ __forceinline__ void rebase_f64_set(const u64 (&v)[E], double (&d)[E], u64 from, const ArF64 &ar)
{
    if ((from >> 51) == 0) {
#pragma unroll
        for (int k = 0; k < E; ++k) d[k] = ar.from_load(v[k]);
        if ((double)from > 2.0 * ar.f.q) {
#pragma unroll
            for (int k = 0; k < E; ++k) d[k] = ar.reduce(d[k]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < E; ++k)
            d[k] = __dadd_rn(ar.mulmod(ar.from_load(v[k] >> 30), 1073741824.0), ar.from_load(v[k] & 0x3FFFFFFFull));
    }
}

// plain transform of `count` limb polynomials [count][N] (measurement API, host tooling)
struct PlainJob {
    static constexpr bool PIPE = true;  // software-pipelined first-pass loads (no epilogue operands -> no spills)
    static constexpr bool R_SMEM = false;
    static constexpr bool INV_OPS = false;  // no epilogue operands on the inverse side
    const u64 *src;
    u64 *dst;
    u32 first_mod, n_mods, n;
    struct R {
        const u64 *src;
        u64 *dst;
    };
    __device__ __forceinline__ u32 mod(u32 j) const { return first_mod + j % n_mods; }
    __device__ __forceinline__ R resolve(u32 j) const { return R{ src + (size_t)j * n, dst + (size_t)j * n }; }
    __device__ __forceinline__ u64 load_raw(const R &r, u32 i, const ModConst &) const { return r.src[i]; }
    __device__ __forceinline__ u64 load_fix(const R &, u64 v, u32, const ModConst &) const { return v; }
    static constexpr bool F64_LOAD = false;
    template <int E, class IdxF>
    __device__ __forceinline__ void fix_set(const R &, u64 (&)[E], IdxF, const ModConst &) const {}
    __device__ __forceinline__ void store(const R &r, u32 i, u64 v, const ModConst &) const { r.dst[i] = v; }
    struct Ops {};
    __device__ __forceinline__ Ops fetch(const R &, u32, const ModConst &) const { return Ops{}; }
    __device__ __forceinline__ void store(const R &r, u32 i, u64 v, const ModConst &m, const Ops &) const { store(r, i, v, m); }
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
    u32 only_c1;      // mod-down of component 1 only (double-hoisted giant steps): job = (e, limb), t is [E][N]
    const u64 *u0;    // ks_inner_sum: accumulators [ngroups*B][2][L+1][N] whose component 0 is gathered through
                      // perm[g] (every limb, also the one mod P) and added to the sum; null: nothing
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
    static constexpr bool PIPE = true;  // gathered loads of the next register set fly during this set's butterflies
    static constexpr bool R_SMEM = true;
    static constexpr bool INV_OPS = false;  // no epilogue operands on the inverse side
    KsParams P;
    struct R {
        const u64 *row;  // limb l of the target polynomial
        const u32 *pm;   // Galois gather table or null
        u64 *dst;
    };
    __device__ __forceinline__ u32 mod(u32 j) const { return j % P.L; }
    __device__ __forceinline__ R resolve(u32 j) const
    {
        const u32 e = j / P.L, l = j % P.L, g = e / P.B, b = e % P.B;
        const CtView &v = P.in[g];
        return R{ v.p + b * v.sb + P.target_poly * v.sp + l * v.sl, P.hoisted ? nullptr : P.perm[g], P.coef + (size_t)j * P.n };
    }
    __device__ __forceinline__ u64 load_raw(const R &r, u32 i, const ModConst &) const { return __ldcg(r.row + (r.pm ? __ldg(r.pm + i) : i)); }
    __device__ __forceinline__ u64 load_fix(const R &, u64 v, u32, const ModConst &) const { return v; }
    __device__ __forceinline__ void store(const R &r, u32 i, u64 x, const ModConst &) const { __stcg(r.dst + i, x); }
};

// K7 step 2: ext[e][j][i] = NTT_{m_i}( c_j mod m_i ), i != j, i in [0, L]
struct KsLiftJob {
    static constexpr bool PIPE = false;
    static constexpr bool R_SMEM = true;
    KsParams P;
    const ModConst *mods;
    struct R {
        const u64 *src;
        u64 *dst;
        u64 from_q;  // modulus of the digit
    };
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
    __device__ __forceinline__ R resolve(u32 j) const
    {
        u32 e, dj, di;
        split(j, e, dj, di);
        return R{ P.coef + ((size_t)e * P.L + dj) * P.n, P.ext + (((size_t)e * P.L + dj) * (P.L + 1) + di) * P.n, mods[dj].q };
    }
    __device__ __forceinline__ u64 load_raw(const R &r, u32 i, const ModConst &) const { return __ldcg(r.src + i); }
    __device__ __forceinline__ u64 load_fix(const R &r, u64 v, u32, const ModConst &m) const { return rebase(v, r.from_q, m); }
    static constexpr bool F64_LOAD = true;
    template <int E, class IdxF>
    __device__ __forceinline__ void fix_set(const R &r, u64 (&v)[E], IdxF, const ModConst &m) const { rebase_set<E>(v, r.from_q, m); }
    template <int E, class IdxF>
    __device__ __forceinline__ void f64_set(const R &r, const u64 (&v)[E], double (&d)[E], IdxF, const ArF64 &ar) const { rebase_f64_set<E>(v, d, r.from_q, ar); }
    struct Ops {};
    __device__ __forceinline__ Ops fetch(const R &, u32, const ModConst &) const { return Ops{}; }
    __device__ __forceinline__ void store(const R &r, u32 i, u64 v, const ModConst &, const Ops &) const { __stcg(r.dst + i, v); }
};

// INTT of a dropped limb with the rounding offset added: t = (INTT_d(src) + floor(d/2)) mod d.
// Used by K7 step 4 (d = special prime) and by rescale (d = q_{L-1}).
struct HalfInttJob {
    static constexpr bool PIPE = false;
    static constexpr bool R_SMEM = false;
    static constexpr bool INV_OPS = false;  // no epilogue operands on the inverse side
    const u64 *src;  // job j at src + (j / inner) * s_outer + (j % inner) * s_inner
    u64 *dst;        // [jobs][N]
    size_t s_outer, s_inner;
    u32 inner, drop_mod, n;
    u32 lazy_in;  // the words are sums of up to 16 canonical residues (multi-GPU partial sums): reduce while loading
    struct R {
        const u64 *src;
        u64 *dst;
    };
    __device__ __forceinline__ u32 mod(u32) const { return drop_mod; }
    __device__ __forceinline__ R resolve(u32 j) const { return R{ src + (j / inner) * s_outer + (j % inner) * s_inner, dst + (size_t)j * n }; }
    __device__ __forceinline__ u64 load_raw(const R &r, u32 i, const ModConst &) const { return r.src[i]; }
    __device__ __forceinline__ u64 load_fix(const R &, u64 v, u32, const ModConst &m) const { return lazy_in ? barrett64(v, m) : v; }
    template <int E, class IdxF>
    __device__ __forceinline__ void fix_set(const R &, u64 (&v)[E], IdxF, const ModConst &m) const
    {
        if (lazy_in) {
#pragma unroll
            for (int k = 0; k < E; ++k) v[k] = barrett64(v[k], m);
        }
    }
    __device__ __forceinline__ void store(const R &r, u32 i, u64 x, const ModConst &m) const { r.dst[i] = addmod(x, m.q >> 1, m.q); }
};

// K7 step 5: out[c][i] = base_c[i] + (acc[c][i] - NTT_{q_i}((t_c mod q_i) - half)) * P^-1
struct KsModDownJob {
    static constexpr bool PIPE = false;
    static constexpr bool R_SMEM = true;
    KsParams P;
    const MdConst *md;  // [K] constants of the dropped modulus (special prime) per target limb
    const ModConst *mods;
    u32 per;  // != 0: jobs in limb-major order, per = jobs per limb (the CTAs that share an SM then run one arithmetic
              // policy, i.e. one third of the kernel's code); 0: limbs interleaved over neighbouring CTAs
    struct R {
        const u64 *t, *acc, *base;  // base: null when nothing is added after the mod-down
        const u32 *pm;              // Galois gather of the base (component 0) or null
        u64 *out;
        u64 from_q, halfmod, inv, inv_sh;
    };
    // job -> (r = polynomial index: e*2 + c, or e when only_c1; l = target limb)
    __device__ __forceinline__ void rl(u32 j, u32 &r, u32 &l) const
    {
        if (per) {
            l = j / per;
            r = j % per;
        } else {
            l = j % P.L;
            r = j / P.L;
        }
    }
    __device__ __forceinline__ u32 mod(u32 j) const
    {
        u32 r, l;
        rl(j, r, l);
        return l;
    }
    __device__ __forceinline__ R resolve(u32 j) const
    {
        u32 rr, l;
        rl(j, rr, l);
        const u32 c = P.only_c1 ? 1u : rr & 1u, e = P.only_c1 ? rr : rr >> 1;
        const u32 g = e / P.B, b = e % P.B;
        const CtView &vi = P.in[g], &vo = P.out[g];
        R r;
        r.t = P.t + (size_t)rr * P.n;
        r.acc = P.acc + (((size_t)e * 2 + c) * (P.L + 1) + l) * P.n;
        r.base = nullptr;
        r.pm = nullptr;
        if (c == 0) {
            if (!P.no_base0) {
                r.base = vi.p + b * vi.sb + l * vi.sl;
                r.pm = P.perm[g];
            }
        } else if (P.has_base1) {
            r.base = vi.p + b * vi.sb + vi.sp + l * vi.sl;
        }
        r.out = vo.p + b * vo.sb + c * vo.sp + l * vo.sl;
        r.from_q = mods[P.K - 1].q;
        r.halfmod = md[l].halfmod;
        r.inv = md[l].inv;
        r.inv_sh = md[l].inv_sh;
        return r;
    }
    __device__ __forceinline__ u64 load_raw(const R &r, u32 i, const ModConst &) const { return __ldcg(r.t + i); }
    __device__ __forceinline__ u64 load_fix(const R &r, u64 v, u32, const ModConst &m) const { return submod(rebase(v, r.from_q, m), r.halfmod, m.q); }
    static constexpr bool F64_LOAD = true;
    template <int E, class IdxF>
    __device__ __forceinline__ void fix_set(const R &r, u64 (&v)[E], IdxF, const ModConst &m) const
    {
        rebase_set<E>(v, r.from_q, m);
        const u64 hm = r.halfmod;
#pragma unroll
        for (int k = 0; k < E; ++k) v[k] = submod(v[k], hm, m.q);
    }
    template <int E, class IdxF>
    __device__ __forceinline__ void f64_set(const R &r, const u64 (&v)[E], double (&d)[E], IdxF, const ArF64 &ar) const
    {
        rebase_f64_set<E>(v, d, r.from_q, ar);
        const double hm = ar.from_load(r.halfmod);
#pragma unroll
        for (int k = 0; k < E; ++k) d[k] = __dadd_rn(d[k], -hm);
    }
    struct Ops {
        u64 acc, base;
    };
    __device__ __forceinline__ Ops fetch(const R &r, u32 i, const ModConst &) const
    {
        Ops o;
        o.acc = __ldcg(r.acc + i);
        o.base = r.base ? __ldcg(r.base + (r.pm ? __ldg(r.pm + i) : i)) : 0;
        return o;
    }
    __device__ __forceinline__ void store(const R &r, u32 i, u64 x, const ModConst &m, const Ops &o) const
    {
        __stcg(r.out + i, addmod(o.base, mul_shoup(submod(o.acc, x, m.q), r.inv, r.inv_sh, m.q), m.q));
    }
};

// rescale step 2: out[b][p][i] = (a[b][p][i] - NTT_{q_i}((t mod q_i) - half)) * q_last^-1
struct RescaleJob {
    static constexpr bool PIPE = false;
    static constexpr bool R_SMEM = true;
    CtView a, out;
    const u64 *t;       // [B*size][N]
    const MdConst *md;  // constants of dropped modulus q_{L-1} per target limb
    const ModConst *mods;
    u32 size, Lm1, drop_mod, n;  // Lm1 = L-1 target limbs
    u32 lazy_in;                 // `a` holds sums of up to 16 canonical residues: reduced in the epilogue's operand fetch
    struct R {
        const u64 *t, *a;
        u64 *out;
        u64 from_q, halfmod, inv, inv_sh;
    };
    __device__ __forceinline__ u32 mod(u32 j) const { return j % Lm1; }
    __device__ __forceinline__ R resolve(u32 j) const
    {
        const u32 l = j % Lm1, bp = j / Lm1, p = bp % size, b = bp / size;
        return R{ t + (size_t)bp * n, a.p + b * a.sb + p * a.sp + l * a.sl, out.p + b * out.sb + p * out.sp + l * out.sl,
                  mods[drop_mod].q, md[l].halfmod, md[l].inv, md[l].inv_sh };
    }
    __device__ __forceinline__ u64 load_raw(const R &r, u32 i, const ModConst &) const { return __ldcg(r.t + i); }
    __device__ __forceinline__ u64 load_fix(const R &r, u64 v, u32, const ModConst &m) const { return submod(rebase(v, r.from_q, m), r.halfmod, m.q); }
    static constexpr bool F64_LOAD = true;
    template <int E, class IdxF>
    __device__ __forceinline__ void fix_set(const R &r, u64 (&v)[E], IdxF, const ModConst &m) const
    {
        rebase_set<E>(v, r.from_q, m);
        const u64 hm = r.halfmod;
#pragma unroll
        for (int k = 0; k < E; ++k) v[k] = submod(v[k], hm, m.q);
    }
    template <int E, class IdxF>
    __device__ __forceinline__ void f64_set(const R &r, const u64 (&v)[E], double (&d)[E], IdxF, const ArF64 &ar) const
    {
        rebase_f64_set<E>(v, d, r.from_q, ar);
        const double hm = ar.from_load(r.halfmod);
#pragma unroll
        for (int k = 0; k < E; ++k) d[k] = __dadd_rn(d[k], -hm);
    }
    struct Ops {
        u64 av;
    };
    __device__ __forceinline__ Ops fetch(const R &r, u32 i, const ModConst &m) const
    {
        const u64 v = __ldcg(r.a + i);
        return Ops{ lazy_in ? barrett64(v, m) : v };
    }
    __device__ __forceinline__ void store(const R &r, u32 i, u64 x, const ModConst &m, const Ops &o) const
    {
        __stcg(r.out + i, mul_shoup(submod(o.av, x, m.q), r.inv, r.inv_sh, m.q));
    }
};

// ---------------------------------------------------------------------------------------
// Mod-down by P followed by rescale by q_last, as ONE pass (bit-identical to the two steps: every
// operation is exact mod q_i and the NTT is linear, so transforms that the two-step form runs back
// to back are merged or cancelled).  With F = accumulator in the basis q_0..q_{L-1},P (NTT form),
// base = ciphertext added after the mod-down, t = INTT_P(F_P) + floor(P/2) (HalfInttJob):
//   limb L-1:  res_last = base + (F - NTT(d)) P^-1, d = (t mod q_last) - (floor(P/2) mod q_last); the rescale needs
//              INTT(res_last) + floor(q_last/2) = INTT(base + F P^-1) - d P^-1 + floor(q_last/2) =: t2
//              -> one INTT with fused loader/epilogue instead of NTT(d) + INTT(res_last)   (FinalInttJob)
//   limb i<L-1: out = (res_i - NTT(d2)) q_last^-1 with d2 = (t2 mod q_i) - (floor(q_last/2) mod q_i)
//              = (base + F P^-1 - NTT(d P^-1 + d2)) q_last^-1
//              -> one NTT instead of two                                                    (FinalNttJob)
// 4 transforms per polynomial instead of 7 at L = 3.
// ---------------------------------------------------------------------------------------
struct FinalParams {
    const u64 *acc;     // [B][2][L+1][N]
    CtView base;        // level-L ciphertext batch added after the mod-down
    CtView out;         // level L-1 result
    const u64 *t;       // [B*2][N]  INTT_P(F_P) + floor(P/2)
    u64 *t2;            // [B*2][N]  INTT_{q_last}(res_last) + floor(q_last/2)
    const MdConst *mdP; // constants of the special prime per target limb
    const MdConst *mdQ; // constants of q_last per target limb
    const ModConst *mods;
    u32 B, L, K, n;
    u32 has_base0, has_base1;
    u32 per;  // FinalNttJob: != 0: jobs in limb-major order, per = jobs per limb (see KsModDownJob)
};
struct FinalInttJob {
    static constexpr bool PIPE = false;
    static constexpr bool R_SMEM = true;
    FinalParams P;
    struct R {
        const u64 *f, *base, *t;  // base: null when nothing is added
        u64 *t2;
        u64 from_q, invP, invP_sh, halfP;
    };
    __device__ __forceinline__ u32 mod(u32) const { return P.L - 1; }
    __device__ __forceinline__ R resolve(u32 j) const
    {
        const u32 b = j >> 1, c = j & 1u, l = P.L - 1;
        return R{ P.acc + ((size_t)j * (P.L + 1) + l) * P.n,
                  (c ? P.has_base1 : P.has_base0) ? P.base.p + b * P.base.sb + c * P.base.sp + l * P.base.sl : nullptr,
                  P.t + (size_t)j * P.n, P.t2 + (size_t)j * P.n, P.mods[P.K - 1].q, P.mdP[l].inv, P.mdP[l].inv_sh, P.mdP[l].halfmod };
    }
    // loader value = (base + F P^-1)[limb L-1], NTT form: load_raw reads F, the fix functions read the base
    __device__ __forceinline__ u64 load_raw(const R &r, u32 i, const ModConst &) const { return __ldcg(r.f + i); }
    __device__ __forceinline__ u64 load_fix(const R &r, u64 v, u32 i, const ModConst &m) const
    {
        u64 g = mul_shoup(v, r.invP, r.invP_sh, m.q);
        if (r.base) g = addmod(g, __ldcg(r.base + i), m.q);
        return g;
    }
    template <int E, class IdxF>
    __device__ __forceinline__ void fix_set(const R &r, u64 (&v)[E], IdxF idx, const ModConst &m) const
    {
        const u64 ip = r.invP, ips = r.invP_sh;
        if (r.base) {
            u64 w[E];
#pragma unroll
            for (int k = 0; k < E; ++k) w[k] = __ldcg(r.base + idx(k));
#pragma unroll
            for (int k = 0; k < E; ++k) v[k] = addmod(mul_shoup(v[k], ip, ips, m.q), w[k], m.q);
        } else {
#pragma unroll
            for (int k = 0; k < E; ++k) v[k] = mul_shoup(v[k], ip, ips, m.q);
        }
    }
    // epilogue operand of position i (fetched for a whole register set before the first element's arithmetic)
    static constexpr bool INV_OPS = true;
    struct IOps {
        u64 t;
    };
    __device__ __forceinline__ IOps ifetch(const R &r, u32 i, const ModConst &) const { return IOps{ __ldcg(r.t + i) }; }
    __device__ __forceinline__ void store(const R &r, u32 i, u64 x, const ModConst &m, const IOps &o) const
    {
        const u64 d = submod(rebase(o.t, r.from_q, m), r.halfP, m.q);
        const u64 v = submod(x, mul_shoup(d, r.invP, r.invP_sh, m.q), m.q);
        __stcg(r.t2 + i, addmod(v, m.q >> 1, m.q));
    }
    __device__ __forceinline__ void store(const R &r, u32 i, u64 x, const ModConst &m) const { store(r, i, x, m, ifetch(r, i, m)); }
};
struct FinalNttJob {
    static constexpr bool PIPE = false;
    static constexpr bool R_SMEM = true;
    FinalParams P;
    struct R {
        const u64 *t, *t2, *f, *base;  // base: null when nothing is added
        u64 *out;
        u64 qP, qL, invP, invP_sh, halfP, invQ, invQ_sh, halfQ;
    };
    __device__ __forceinline__ void rl(u32 j, u32 &bc, u32 &l) const
    {
        if (P.per) {
            l = j / P.per;
            bc = j % P.per;
        } else {
            l = j % (P.L - 1);
            bc = j / (P.L - 1);
        }
    }
    __device__ __forceinline__ u32 mod(u32 j) const
    {
        u32 bc, l;
        rl(j, bc, l);
        return l;
    }
    __device__ __forceinline__ R resolve(u32 j) const
    {
        u32 bc, l;
        rl(j, bc, l);
        const u32 b = bc >> 1, c = bc & 1u;
        return R{ P.t + (size_t)bc * P.n, P.t2 + (size_t)bc * P.n, P.acc + ((size_t)bc * (P.L + 1) + l) * P.n,
                  (c ? P.has_base1 : P.has_base0) ? P.base.p + b * P.base.sb + c * P.base.sp + l * P.base.sl : nullptr,
                  P.out.p + b * P.out.sb + c * P.out.sp + l * P.out.sl, P.mods[P.K - 1].q, P.mods[P.L - 1].q,
                  P.mdP[l].inv, P.mdP[l].inv_sh, P.mdP[l].halfmod, P.mdQ[l].inv, P.mdQ[l].inv_sh, P.mdQ[l].halfmod };
    }
    // loader value = d P^-1 + d2 (coefficient form), d = (t mod q) - (floor(P/2) mod q), d2 = (t2 mod q) - (floor(q_last/2) mod q):
    // load_raw reads t, the fix functions read t2 (all loads of a register set are issued before its arithmetic)
    __device__ __forceinline__ u64 load_raw(const R &r, u32 i, const ModConst &) const { return __ldcg(r.t + i); }
    __device__ __forceinline__ u64 load_fix(const R &r, u64 v, u32 i, const ModConst &m) const
    {
        const u64 d = submod(rebase(v, r.qP, m), r.halfP, m.q);
        const u64 d2 = submod(rebase(__ldcg(r.t2 + i), r.qL, m), r.halfQ, m.q);
        return addmod(mul_shoup(d, r.invP, r.invP_sh, m.q), d2, m.q);
    }
    static constexpr bool F64_LOAD = true;
    template <int E, class IdxF>
    __device__ __forceinline__ void fix_set(const R &r, u64 (&v)[E], IdxF idx, const ModConst &m) const
    {
        u64 w[E];
#pragma unroll
        for (int k = 0; k < E; ++k) w[k] = __ldcg(r.t2 + idx(k));
        rebase_set<E>(v, r.qP, m);
        rebase_set<E>(w, r.qL, m);
        const u64 hp = r.halfP, hq = r.halfQ, ip = r.invP, ips = r.invP_sh;
#pragma unroll
        for (int k = 0; k < E; ++k)
            v[k] = addmod(mul_shoup(submod(v[k], hp, m.q), ip, ips, m.q), submod(w[k], hq, m.q), m.q);
    }
    template <int E, class IdxF>
    __device__ __forceinline__ void f64_set(const R &r, const u64 (&v)[E], double (&d)[E], IdxF idx, const ArF64 &ar) const
    {
        u64 w[E];
#pragma unroll
        for (int k = 0; k < E; ++k) w[k] = __ldcg(r.t2 + idx(k));
        double d2[E];
        rebase_f64_set<E>(v, d, r.qP, ar);
        rebase_f64_set<E>(w, d2, r.qL, ar);
        const double hp = ar.from_load(r.halfP), hq = ar.from_load(r.halfQ), ip = ar.from_load(r.invP);
#pragma unroll
        for (int k = 0; k < E; ++k) d[k] = __dadd_rn(ar.mulmod(__dadd_rn(d[k], -hp), ip), __dadd_rn(d2[k], -hq));
    }
    struct Ops {
        u64 f, base;
    };
    __device__ __forceinline__ Ops fetch(const R &r, u32 i, const ModConst &) const { return Ops{ __ldcg(r.f + i), r.base ? __ldcg(r.base + i) : 0 }; }
    __device__ __forceinline__ void store(const R &r, u32 i, u64 x, const ModConst &m, const Ops &o) const
    {
        const u64 v = addmod(o.base, mul_shoup(o.f, r.invP, r.invP_sh, m.q), m.q);
        __stcg(r.out + i, mul_shoup(submod(v, x, m.q), r.invQ, r.invQ_sh, m.q));
    }
};

// ---------------------------------------------------------------------------------------
// Loaders (protocol in ntt.cuh): raw() = memory only, fix() = arithmetic only.
// ---------------------------------------------------------------------------------------
// rows and constants of this CTA's job (Job::R): in registers when small, else computed by thread 0 into shared memory
#define HEGPU_RESOLVE_JOB(job, jid)                          \
    __shared__ typename Job::R r_sh;                         \
    typename Job::R r_loc;                                   \
    if constexpr (Job::R_SMEM) {                             \
        if (threadIdx.x == 0) r_sh = (job).resolve(jid);     \
        __syncthreads();                                     \
    } else {                                                 \
        r_loc = (job).resolve(jid);                          \
    }                                                        \
    const typename Job::R &r = *(Job::R_SMEM ? &r_sh : &r_loc)

template <class Job>
struct PlainLoader {  // coefficient boff + i of the resolved job
    typedef u64 Raw;
    static constexpr bool PIPE = Job::PIPE;
    const Job &job;
    const ModConst &m;
    const typename Job::R &jr;
    u32 boff;
    __device__ __forceinline__ Raw raw(u32 i) const { return job.load_raw(jr, boff + i, m); }
    __device__ __forceinline__ u64 fix(Raw r, u32 i) const { return job.load_fix(jr, r, boff + i, m); }
    // a whole register set at once (forward pass 0): idx(k) = coefficient index of element k
    template <int E, class A, class IdxF>
    __device__ __forceinline__ void fix_set(const Raw (&raw)[E], typename A::V (&x)[E], IdxF idx, const A &ar) const
    {
        u64 v[E];
#pragma unroll
        for (int k = 0; k < E; ++k) v[k] = raw[k];
        job.template fix_set<E>(jr, v, [&](int k) { return boff + idx(k); }, m);
#pragma unroll
        for (int k = 0; k < E; ++k) x[k] = ar.from_load(v[k]);
    }
};
// product of the stride-N/2 stage that the fold loaders compute while loading (canonical inputs): with the approximate
// quotient the outputs are below 4q instead of 3q, which the range analysis of ArI64 covers (pass 0 has no correction and
// ends below 4q + 8q)
__device__ __forceinline__ u64 fold_mul(u64 y, const ulonglong2 W, const ModConst &m)
{
    return HEGPU_SHOUP_APPROX ? mul_shoup_lazy3_nq(y, W.x, W.y, 0ull - m.q) : mul_shoup_lazy(y, W.x, W.y, m.q);
}
__device__ __forceinline__ u64 fold_off(const ModConst &m) { return HEGPU_SHOUP_APPROX ? m.q3 : m.q << 1; }
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
    const typename Job::R &jr;
    u32 h;
    ulonglong2 W;  // twiddle of the first stage
    u64 *park;
    __device__ __forceinline__ Raw raw(u32 i) const { return Pair64{ job.load_raw(jr, i, m), job.load_raw(jr, i + (1u << LOGL), m) }; }
    __device__ __forceinline__ u64 fix(Raw r, u32 i) const
    {
        const u64 X = job.load_fix(jr, r.x, i, m), Y = job.load_fix(jr, r.y, i + (1u << LOGL), m);
        const u64 Tm = fold_mul(Y, W, m);
        const u64 top = X + Tm, bot = X + fold_off(m) - Tm;
        if (park) park[i] = h ? top : bot;
        return h ? bot : top;
    }
    template <int E, class A, class IdxF>
    __device__ __forceinline__ void fix_set(const Raw (&raw)[E], typename A::V (&x)[E], IdxF idx, const A &ar) const
    {
#pragma unroll
        for (int k = 0; k < E; ++k) x[k] = ar.from_load(fix(raw[k], idx(k)));
    }
};
// loader of the park kernels: half 0 folds the stride-N/2 stage from the job's input and parks the
// bottom outputs; half 1 reads what half 0 parked (same thread wrote it).  One type for both halves
// so that the two passes of a park kernel share their code (the instruction cache is the limit:
// ncu showed 19 % of the stall samples of the duplicated version as stall_no_inst).
template <class Job, int LOGL>
struct ParkFoldLoader {
    typedef Pair64 Raw;
    static constexpr bool PIPE = Job::PIPE;
    const Job &job;
    const ModConst &m;
    const typename Job::R &jr;
    u32 half;
    ulonglong2 W;  // twiddle of the first stage
    double Wd;     // the same as a double (FP64 policy)
    u64 *park;
    __device__ __forceinline__ Raw raw(u32 i) const
    {
        if (half) return Pair64{ park[i], 0 };
        return Pair64{ job.load_raw(jr, i, m), job.load_raw(jr, i + (1u << LOGL), m) };
    }
    __device__ __forceinline__ u64 fix(Raw r, u32 i) const
    {
        if (half) return r.x;
        const u64 X = job.load_fix(jr, r.x, i, m), Y = job.load_fix(jr, r.y, i + (1u << LOGL), m);
        const u64 Tm = fold_mul(Y, W, m);
        park[i] = X + fold_off(m) - Tm;
        return X + Tm;
    }
    // A whole register set at once: the branches on the half and on the pair of moduli are taken once per set.  In the FP64
    // policy (jobs with F64_LOAD) the operands become doubles straight away and the folded stage runs on the FP64 pipe; the
    // parked half then holds doubles.
    template <int E, class A, class IdxF>
    __device__ __forceinline__ void fix_set(const Raw (&raw)[E], typename A::V (&x)[E], IdxF idx, const A &ar) const
    {
        constexpr bool F64 = std::is_same<A, ArF64>::value && Job::F64_LOAD;
        if (half) {
#pragma unroll
            for (int k = 0; k < E; ++k) {
                if constexpr (F64)
                    x[k] = __longlong_as_double((long long)raw[k].x);
                else
                    x[k] = ar.from_load(raw[k].x);
            }
            return;
        }
        u64 X[E], Y[E];
#pragma unroll
        for (int k = 0; k < E; ++k) {
            X[k] = raw[k].x;
            Y[k] = raw[k].y;
        }
        if constexpr (F64) {
            double dx[E], dy[E];
            job.template f64_set<E>(jr, X, dx, idx, ar);
            job.template f64_set<E>(jr, Y, dy, [&](int k) { return idx(k) + (1u << LOGL); }, ar);
#pragma unroll
            for (int k = 0; k < E; ++k) {
                const double Tm = ar.mulmod(dy[k], Wd);
                park[idx(k)] = (u64)__double_as_longlong(__dadd_rn(dx[k], -Tm));
                x[k] = __dadd_rn(dx[k], Tm);
            }
        } else {
            job.template fix_set<E>(jr, X, idx, m);
            job.template fix_set<E>(jr, Y, [&](int k) { return idx(k) + (1u << LOGL); }, m);
            const u64 off = fold_off(m);
#pragma unroll
            for (int k = 0; k < E; ++k) {
                const u64 Tm = fold_mul(Y[k], W, m);
                park[idx(k)] = X[k] + off - Tm;
                x[k] = ar.from_load(X[k] + Tm);
            }
        }
    }
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
    HEGPU_RESOLVE_JOB(job, jid);
    const ModConst m = T.mods[mi];
    const ulonglong2 *tw = T.fwd + (size_t)mi * T.n;
    const u32 boff = h << LOGL;
    auto fetch = [&](u32 i) { return job.fetch(r, boff + i, m); };
    auto store = [&](u32 i, u64 v, const typename Job::Ops &o) { job.store(r, boff + i, v, m, o); };
    const ulonglong2 nowl = make_ulonglong2(0, 0);
    if constexpr (SPLIT == 0) {
        PlainLoader<Job> load{ job, m, r, 0 };
        if (m.big & 4u)
            ntt_fwd_cta<LOGL, LOGE>(load, fetch, store, T.fwd_d + (size_t)mi * T.n, T.n, ArF64(T.modsd[mi]), sm);
        else if (m.big & 1u)
            ntt_fwd_cta<LOGL, LOGE>(load, fetch, store, tw, T.n, ArI64<true>(m, nowl), sm);
        else
            ntt_fwd_cta<LOGL, LOGE>(load, fetch, store, tw, T.n, ArI64<false>(m, nowl), sm);
    } else {
        // stage 1 (stride N/2) redone from global memory by both halves
        FoldLoader<Job, LOGL> load{ job, m, r, h, __ldg(tw + 1), nullptr };
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
    HEGPU_RESOLVE_JOB(job, jid);
    const ModConst m = T.mods[mi];
    const ulonglong2 *tw = T.inv + (size_t)mi * T.n;
    const ulonglong2 wl = T.inv_last[mi];
    const u32 boff = h << LOGL;
    PlainLoader<Job> load{ job, m, r, boff };
    if constexpr (SPLIT == 0) {
        auto store = [&](const u32 (&idx)[1 << LOGE], const u64 (&v)[1 << LOGE]) {
#pragma unroll
            for (int k = 0; k < (1 << LOGE); ++k) job.store(r, idx[k], v[k], m);
        };
        if (m.big & 4u)
            ntt_inv_cta<LOGL, LOGE, LOGL - 1>(load, store, T.inv_d + (size_t)mi * T.n, T.n, ArF64(T.modsd[mi]), sm);
        else if (m.big & 2u)
            ntt_inv_cta<LOGL, LOGE, LOGL - 1>(load, store, tw, T.n, ArI64<true>(m, wl), sm);
        else
            ntt_inv_cta<LOGL, LOGE, LOGL - 1>(load, store, tw, T.n, ArI64<false>(m, wl), sm);
    } else {
        u64 *dst = scratch + (size_t)jid * T.n + boff;
        auto store = [&](const u32 (&idx)[1 << LOGE], const u64 (&v)[1 << LOGE]) {
#pragma unroll
            for (int k = 0; k < (1 << LOGE); ++k) dst[idx[k]] = v[k];
        };
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
    HEGPU_RESOLVE_JOB(job, jid);
    const ModConst m = T.mods[mi];
    const ulonglong2 *tw = T.fwd + (size_t)mi * T.n;
    constexpr u32 half = 1u << LOGL;
    u64 *pk = park + (size_t)jid * half;
    const double *twd0 = T.fwd_d + (size_t)mi * T.n;
    ParkFoldLoader<Job, LOGL> load{ job, m, r, 0, __ldg(tw + 1), (m.big & 4u) ? __ldg(twd0 + 1) : 0.0, pk };
    u32 boff = 0;  // block offset of the half being transformed
    auto fetch = [&](u32 i) { return job.fetch(r, boff + i, m); };
    auto store = [&](u32 i, u64 v, const typename Job::Ops &o) { job.store(r, boff + i, v, m, o); };
    const ulonglong2 nowl = make_ulonglong2(0, 0);
    if (m.big & 4u) {
        const ArF64 ar(T.modsd[mi]);
        const double *twd = T.fwd_d + (size_t)mi * T.n;
#pragma unroll 1
        for (u32 h = 0; h < 2; ++h) {
            load.half = h;
            boff = h << LOGL;
            ntt_fwd_cta<LOGL, LOGE>(load, fetch, store, twd, T.n + boff, ar, sm);
            __syncthreads();
        }
    } else if (m.big & 1u) {
        const ArI64<true> ar(m, nowl);
#pragma unroll 1
        for (u32 h = 0; h < 2; ++h) {
            load.half = h;
            boff = h << LOGL;
            ntt_fwd_cta<LOGL, LOGE>(load, fetch, store, tw, T.n + boff, ar, sm);
            __syncthreads();
        }
    } else {
        const ArI64<false> ar(m, nowl);
#pragma unroll 1
        for (u32 h = 0; h < 2; ++h) {
            load.half = h;
            boff = h << LOGL;
            ntt_fwd_cta<LOGL, LOGE>(load, fetch, store, tw, T.n + boff, ar, sm);
            __syncthreads();
        }
    }
}

template <int LOGL, int LOGE, class A, class TWP, class Job>
__device__ __forceinline__ void inv_park_body(const Job &job, const typename Job::R &r, const ModConst &m, const A &ar, const TWP tw, u32 n,
                                              typename A::V *pk, u64 *sm)
{
    constexpr u32 half = 1u << LOGL;
    PlainLoader<Job> load{ job, m, r, 0 };
    u32 h = 0;
    // half 0 is transformed and parked; half 1 is transformed and its last-pass registers are combined
    // with the parked half in the final stride-N/2 stage.  One store for both halves: shared code.
    constexpr int E = 1 << LOGE;
    auto store = [&](const u32 (&idx)[E], typename A::V (&y)[E]) {
        if (h == 0) {
#pragma unroll
            for (int k = 0; k < E; ++k) pk[idx[k]] = y[k];
            return;
        }
        if constexpr (Job::INV_OPS) {
            // half a register set at a time: parked partner and the two epilogue operands of every element first
            constexpr int H = E / 2;
#pragma unroll
            for (int k0 = 0; k0 < E; k0 += H) {
                typename A::V x[H];
                typename Job::IOps o0[H], o1[H];
#pragma unroll
                for (int k = 0; k < H; ++k) {
                    x[k] = __ldcg(pk + idx[k0 + k]);
                    o0[k] = job.ifetch(r, idx[k0 + k], m);
                    o1[k] = job.ifetch(r, half + idx[k0 + k], m);
                }
#pragma unroll
                for (int k = 0; k < H; ++k) {
                    ar.template inv_bfly_last<LOGL>(x[k], y[k0 + k]);
                    job.store(r, idx[k0 + k], ar.inv_final(x[k]), m, o0[k]);
                    job.store(r, half + idx[k0 + k], ar.inv_final(y[k0 + k]), m, o1[k]);
                }
            }
        } else {
            typename A::V x[E];  // all parked loads first: one L2 latency per register set, not one per element
#pragma unroll
            for (int k = 0; k < E; ++k) x[k] = __ldcg(pk + idx[k]);
#pragma unroll
            for (int k = 0; k < E; ++k) {
                ar.template inv_bfly_last<LOGL>(x[k], y[k]);
                job.store(r, idx[k], ar.inv_final(x[k]), m);
                job.store(r, half + idx[k], ar.inv_final(y[k]), m);
            }
        }
    };
#pragma unroll 1
    for (h = 0; h < 2; ++h) {
        load.boff = h << LOGL;
        ntt_inv_cta<LOGL, LOGE, -1>(load, store, tw, n + (h << LOGL), ar, sm);
        __syncthreads();
    }
}

template <int LOGL, int LOGE, class Job>
__global__ void __launch_bounds__(NttShape<LOGL, LOGE>::THREADS, NttShape<LOGL, LOGE>::MINB)
    ntt_inv_park_kernel(const Job job, const NttTables T, u64 *__restrict__ park)
{
    extern __shared__ __align__(16) u64 sm[];
    const u32 jid = blockIdx.x;
    const u32 mi = job.mod(jid);
    HEGPU_RESOLVE_JOB(job, jid);
    const ModConst m = T.mods[mi];
    const ulonglong2 *tw = T.inv + (size_t)mi * T.n;
    u64 *pk = park + ((size_t)jid << LOGL);
    if (m.big & 4u)
        inv_park_body<LOGL, LOGE>(job, r, m, ArF64(T.modsd[mi]), T.inv_d + (size_t)mi * T.n, T.n, reinterpret_cast<double *>(pk), sm);
    else if (m.big & 2u)
        inv_park_body<LOGL, LOGE>(job, r, m, ArI64<true>(m, T.inv_last[mi]), tw, T.n, pk, sm);
    else
        inv_park_body<LOGL, LOGE>(job, r, m, ArI64<false>(m, T.inv_last[mi]), tw, T.n, pk, sm);
}

// ---------------------------------------------------------------------------------------
// Two-level park kernels: N = 2^(LOGL+2) (N = 32768 at LOGL = 13) transformed by ONE CTA in 2^LOGL words of shared
// memory, so that N = 32768 runs the very same 64 KiB / 256-thread / three-CTAs-per-SM sub-transforms as N = 16384
// (the one-level form needs 128 KiB and leaves one CTA per SM, whose load, exchange and store phases cannot overlap).
// Forward: the stride-N/2 and stride-N/4 stages (a radix-4 butterfly on x[i], x[i+N/4], x[i+N/2], x[i+3N/4]) are
// computed while loading quarter 0; the other three outputs are parked ([jobs][3][N/4] words, written and re-read by
// the same thread) and transformed as quarters 1..3.  Inverse: quarters 0..2 are transformed and parked, quarter 3 is
// transformed and its last-pass registers are combined with the parked quarters in the final two stages.
// ---------------------------------------------------------------------------------------
struct Quad64 {
    u64 a, b, c, d;
};
template <class Job, int LOGL>
struct Park4FoldLoader {
    typedef Quad64 Raw;
    static constexpr bool PIPE = false;
    const Job &job;
    const ModConst &m;
    const typename Job::R &jr;
    u32 quarter;
    ulonglong2 W1, W2, W3;  // twiddles of the first stage (index 1) and of the second (indices 2, 3)
    double W1d, W2d, W3d;   // the same as doubles (FP64 policy)
    u64 *park;              // [3][2^LOGL]
    __device__ __forceinline__ Raw raw(u32 i) const
    {
        constexpr u32 Q = 1u << LOGL;
        if (quarter) return Quad64{ park[(quarter - 1) * Q + i], 0, 0, 0 };
        return Quad64{ job.load_raw(jr, i, m), job.load_raw(jr, i + Q, m), job.load_raw(jr, i + 2 * Q, m), job.load_raw(jr, i + 3 * Q, m) };
    }
    // canonical inputs; outputs below 7q with the approximate quotient (q -> 4q -> 7q), 5q with the exact one: pass 0 of
    // the sub-transform has no correction and ends below 7q + 8q < 16q
    __device__ __forceinline__ u64 fix(Raw r, u32 i) const
    {
        constexpr u32 Q = 1u << LOGL;
        if (quarter) return r.a;
        const u64 a = job.load_fix(jr, r.a, i, m), b = job.load_fix(jr, r.b, i + Q, m), c = job.load_fix(jr, r.c, i + 2 * Q, m), d = job.load_fix(jr, r.d, i + 3 * Q, m);
        const u64 off = fold_off(m);
        u64 T = fold_mul(c, W1, m);
        const u64 a1 = a + T, c1 = a + off - T;
        T = fold_mul(d, W1, m);
        const u64 b1 = b + T, d1 = b + off - T;
        T = fold_mul(b1, W2, m);
        park[i] = a1 + off - T;
        const u64 a2 = a1 + T;
        T = fold_mul(d1, W3, m);
        park[Q + i] = c1 + T;
        park[2 * Q + i] = c1 + off - T;
        return a2;
    }
    template <int E, class A, class IdxF>
    __device__ __forceinline__ void fix_set(const Raw (&raw)[E], typename A::V (&x)[E], IdxF idx, const A &ar) const
    {
        constexpr u32 Q = 1u << LOGL;
        constexpr bool F64 = std::is_same<A, ArF64>::value && Job::F64_LOAD;
        if (quarter) {
#pragma unroll
            for (int k = 0; k < E; ++k) {
                if constexpr (F64)
                    x[k] = __longlong_as_double((long long)raw[k].a);
                else
                    x[k] = ar.from_load(raw[k].a);
            }
            return;
        }
        u64 va[E], vb[E], vc[E], vd[E];
#pragma unroll
        for (int k = 0; k < E; ++k) {
            va[k] = raw[k].a;
            vb[k] = raw[k].b;
            vc[k] = raw[k].c;
            vd[k] = raw[k].d;
        }
        if constexpr (F64) {
            double a[E], b[E], c[E], d[E];
            job.template f64_set<E>(jr, va, a, idx, ar);
            job.template f64_set<E>(jr, vb, b, [&](int k) { return idx(k) + Q; }, ar);
            job.template f64_set<E>(jr, vc, c, [&](int k) { return idx(k) + 2 * Q; }, ar);
            job.template f64_set<E>(jr, vd, d, [&](int k) { return idx(k) + 3 * Q; }, ar);
#pragma unroll
            for (int k = 0; k < E; ++k) {
                const u32 i = idx(k);
                double T = ar.mulmod(c[k], W1d);
                const double a1 = __dadd_rn(a[k], T), c1 = __dadd_rn(a[k], -T);
                T = ar.mulmod(d[k], W1d);
                const double b1 = __dadd_rn(b[k], T), d1 = __dadd_rn(b[k], -T);
                T = ar.mulmod(b1, W2d);
                park[i] = (u64)__double_as_longlong(__dadd_rn(a1, -T));
                x[k] = __dadd_rn(a1, T);
                T = ar.mulmod(d1, W3d);
                park[Q + i] = (u64)__double_as_longlong(__dadd_rn(c1, T));
                park[2 * Q + i] = (u64)__double_as_longlong(__dadd_rn(c1, -T));
            }
        } else {
            job.template fix_set<E>(jr, va, idx, m);
            job.template fix_set<E>(jr, vb, [&](int k) { return idx(k) + Q; }, m);
            job.template fix_set<E>(jr, vc, [&](int k) { return idx(k) + 2 * Q; }, m);
            job.template fix_set<E>(jr, vd, [&](int k) { return idx(k) + 3 * Q; }, m);
            const u64 off = fold_off(m);
#pragma unroll
            for (int k = 0; k < E; ++k) {
                const u32 i = idx(k);
                u64 T = fold_mul(vc[k], W1, m);
                const u64 a1 = va[k] + T, c1 = va[k] + off - T;
                T = fold_mul(vd[k], W1, m);
                const u64 b1 = vb[k] + T, d1 = vb[k] + off - T;
                T = fold_mul(b1, W2, m);
                park[i] = a1 + off - T;
                x[k] = ar.from_load(a1 + T);
                T = fold_mul(d1, W3, m);
                park[Q + i] = c1 + T;
                park[2 * Q + i] = c1 + off - T;
            }
        }
    }
};

template <int LOGL, int LOGE, class Job>
__global__ void __launch_bounds__(NttShape<LOGL, LOGE>::THREADS, NttShape<LOGL, LOGE>::MINB)
    ntt_fwd_park4_kernel(const Job job, const NttTables T, u64 *__restrict__ park)
{
    extern __shared__ __align__(16) u64 sm[];
    const u32 jid = blockIdx.x;
    const u32 mi = job.mod(jid);
    HEGPU_RESOLVE_JOB(job, jid);
    const ModConst m = T.mods[mi];
    const ulonglong2 *tw = T.fwd + (size_t)mi * T.n;
    u64 *pk = park + (size_t)jid * (3u << LOGL);
    const double *twd0 = T.fwd_d + (size_t)mi * T.n;
    const bool f64 = (m.big & 4u) != 0;
    Park4FoldLoader<Job, LOGL> load{ job, m, r, 0, __ldg(tw + 1), __ldg(tw + 2), __ldg(tw + 3),
                                     f64 ? __ldg(twd0 + 1) : 0.0, f64 ? __ldg(twd0 + 2) : 0.0, f64 ? __ldg(twd0 + 3) : 0.0, pk };
    u32 boff = 0;  // block offset of the quarter being transformed
    auto fetch = [&](u32 i) { return job.fetch(r, boff + i, m); };
    auto store = [&](u32 i, u64 v, const typename Job::Ops &o) { job.store(r, boff + i, v, m, o); };
    const ulonglong2 nowl = make_ulonglong2(0, 0);
    if (m.big & 4u) {
        const ArF64 ar(T.modsd[mi]);
        const double *twd = T.fwd_d + (size_t)mi * T.n;
#pragma unroll 1
        for (u32 h = 0; h < 4; ++h) {
            load.quarter = h;
            boff = h << LOGL;
            ntt_fwd_cta<LOGL, LOGE>(load, fetch, store, twd, T.n + boff, ar, sm);
            __syncthreads();
        }
    } else if (m.big & 1u) {
        const ArI64<true> ar(m, nowl);
#pragma unroll 1
        for (u32 h = 0; h < 4; ++h) {
            load.quarter = h;
            boff = h << LOGL;
            ntt_fwd_cta<LOGL, LOGE>(load, fetch, store, tw, T.n + boff, ar, sm);
            __syncthreads();
        }
    } else {
        const ArI64<false> ar(m, nowl);
#pragma unroll 1
        for (u32 h = 0; h < 4; ++h) {
            load.quarter = h;
            boff = h << LOGL;
            ntt_fwd_cta<LOGL, LOGE>(load, fetch, store, tw, T.n + boff, ar, sm);
            __syncthreads();
        }
    }
}

template <int LOGL, int LOGE, class A, class TWP, class Job>
__device__ __forceinline__ void inv_park4_body(const Job &job, const typename Job::R &r, const ModConst &m, const A &ar, const TWP tw, u32 n,
                                               typename A::V *pk, u64 *sm)
{
    constexpr u32 Q = 1u << LOGL;
    PlainLoader<Job> load{ job, m, r, 0 };
    u32 h = 0;
    const typename A::TW W2 = __ldg(tw + 2), W3 = __ldg(tw + 3);  // stage LOGL: quarters (0,1) and (2,3)
    constexpr int E = 1 << LOGE;
    auto store = [&](const u32 (&idx)[E], typename A::V (&y)[E]) {
        if (h < 3) {
#pragma unroll
            for (int k = 0; k < E; ++k) pk[h * Q + idx[k]] = y[k];
            return;
        }
        // parked loads of half a register set at a time (three per element: a whole set would not fit the register budget)
        constexpr int H = E / 2;
#pragma unroll
        for (int k0 = 0; k0 < E; k0 += H) {
            typename A::V x0[H], x1[H], x2[H];
#pragma unroll
            for (int k = 0; k < H; ++k) {
                x0[k] = __ldcg(pk + idx[k0 + k]);
                x1[k] = __ldcg(pk + Q + idx[k0 + k]);
                x2[k] = __ldcg(pk + 2 * Q + idx[k0 + k]);
            }
#pragma unroll
            for (int k = 0; k < H; ++k) {
                const u32 i = idx[k0 + k];
                ar.template inv_bfly<LOGL>(x0[k], x1[k], W2);
                ar.template inv_bfly<LOGL>(x2[k], y[k0 + k], W3);
                ar.template inv_bfly_last<LOGL + 1>(x0[k], x2[k]);
                ar.template inv_bfly_last<LOGL + 1>(x1[k], y[k0 + k]);
                job.store(r, i, ar.inv_final(x0[k]), m);
                job.store(r, Q + i, ar.inv_final(x1[k]), m);
                job.store(r, 2 * Q + i, ar.inv_final(x2[k]), m);
                job.store(r, 3 * Q + i, ar.inv_final(y[k0 + k]), m);
            }
        }
    };
#pragma unroll 1
    for (h = 0; h < 4; ++h) {
        load.boff = h << LOGL;
        ntt_inv_cta<LOGL, LOGE, -1>(load, store, tw, n + (h << LOGL), ar, sm);
        __syncthreads();
    }
}

template <int LOGL, int LOGE, class Job>
__global__ void __launch_bounds__(NttShape<LOGL, LOGE>::THREADS, NttShape<LOGL, LOGE>::MINB)
    ntt_inv_park4_kernel(const Job job, const NttTables T, u64 *__restrict__ park)
{
    extern __shared__ __align__(16) u64 sm[];
    const u32 jid = blockIdx.x;
    const u32 mi = job.mod(jid);
    HEGPU_RESOLVE_JOB(job, jid);
    const ModConst m = T.mods[mi];
    const ulonglong2 *tw = T.inv + (size_t)mi * T.n;
    u64 *pk = park + (size_t)jid * (3u << LOGL);
    if (m.big & 4u)
        inv_park4_body<LOGL, LOGE>(job, r, m, ArF64(T.modsd[mi]), T.inv_d + (size_t)mi * T.n, T.n, reinterpret_cast<double *>(pk), sm);
    else if (m.big & 2u)
        inv_park4_body<LOGL, LOGE>(job, r, m, ArI64<true>(m, T.inv_last[mi]), tw, T.n, pk, sm);
    else
        inv_park4_body<LOGL, LOGE>(job, r, m, ArI64<false>(m, T.inv_last[mi]), tw, T.n, pk, sm);
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
        const typename Job::R r = job.resolve(jid);
        const ModConst m = T.mods[mi];
        const ulonglong2 wl = T.inv_last[mi];
        const u64 X = scratch[(size_t)jid * T.n + i], Y = scratch[(size_t)jid * T.n + halfn + i];
        // lazy ranges of ArI64::inv_bfly: [0, LZ q) (corrected per stage) or [0, LZ q * 2^LOGL) (sums left to double)
        const u64 Sm = X + Y;
        const u64 D = (m.big & 2u) ? X + (ArI64<true>::LZ == 3 ? m.q3 : m.q << 1) - Y : X + (m.q << (LOGL + 2)) - Y;
        job.store(r, i, mul_shoup(Sm, m.ninv, m.ninv_sh, m.q), m);
        job.store(r, i + halfn, mul_shoup(D, wl.x, wl.y, m.q), m);
    }
}

}  // namespace hegpu
