// modarith.cuh -- 64-bit modular arithmetic for sm_100a, built on 32-bit IMAD chains.
// Canonical residues in [0,q), q < 2^61.  Constants (Shoup quotients, Barrett ratios) are
// precomputed on the host (context.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hegpu {

typedef unsigned long long u64;
typedef unsigned int u32;

// per-modulus constants, one per prime of the key-level chain
struct ModConst {
    u64 q;
    u64 mu_hi;    // floor(2^128 / q) >> 64  == floor(2^64 / q)
    u64 mu_lo;    // floor(2^128 / q) low word
    u64 ninv;     // N^{-1} mod q
    u64 ninv_sh;  // Shoup quotient of ninv
    u32 big;      // bit 0/1: forward / inverse NTT needs lazy-range corrections; bit 2: FP64 butterflies
    u32 pad;
    u64 qinv_neg; // -q^-1 mod 2^64 (Montgomery reduction)
    u64 rmod;     // 2^64 mod q, and its Shoup quotient: x -> x*2^64 mod q (Montgomery form)
    u64 rmod_sh;
    u64 pmont;    // (special prime P mod q) * 2^64 mod q: x -> x*P mod q through one Montgomery reduction
    u64 q3;       // 3q, loaded rather than computed: ptxas otherwise rebuilds it with an IMAD.WIDE + IMAD in every butterfly
};

__device__ __forceinline__ u64 csub(u64 v, u64 c) { return v >= c ? v - c : v; }

// x * w mod q, lazily in [0, 2q); wsh = floor(w * 2^64 / q); any x < 2^64; nq = 2^64 - q, so that x*w + h*nq is one
// multiply-add chain (no negation of h*q).  Written as 32-bit carry chains: ptxas emits 4 IMAD.WIDE + 2 IMAD.HI + 4 IMAD
// and keeps the additions on the ALU pipe; the C form (__umul64hi) compiles to 6 IMAD.WIDE + 4 IMAD plus ~1.5
// IMAD.X / IMAD.IADD / IMAD.MOV per product on the multiplier pipe that bounds the 60-bit transforms
// (tools/bfly_variants.cu: 33.6 instead of 35.6 multiplier-pipe cycles per butterfly).
__device__ __forceinline__ u64 mul_shoup_lazy_nq(u64 x, u64 w, u64 wsh, u64 nq)
{
    u64 r;
    asm("{\n\t"
        ".reg .u32 x0, x1, w0, w1, s0, s1, n0, n1, c0, m0, m1, m2, h0, h1, r0, r1;\n\t"
        "mov.b64 {x0, x1}, %1;\n\t"
        "mov.b64 {w0, w1}, %2;\n\t"
        "mov.b64 {s0, s1}, %3;\n\t"
        "mov.b64 {n0, n1}, %4;\n\t"
        "mul.hi.u32     c0, x0, s0;\n\t"
        "mad.lo.cc.u32  m0, x0, s1, c0;\n\t"
        "madc.hi.u32    m1, x0, s1, 0;\n\t"
        "mad.lo.cc.u32  m0, x1, s0, m0;\n\t"
        "madc.hi.cc.u32 m1, x1, s0, m1;\n\t"
        "addc.u32       m2, 0, 0;\n\t"
        "mad.lo.cc.u32  h0, x1, s1, m1;\n\t"
        "madc.hi.u32    h1, x1, s1, m2;\n\t"
        "mul.lo.u32     r0, x0, w0;\n\t"
        "mul.hi.u32     r1, x0, w0;\n\t"
        "mad.lo.cc.u32  r0, h0, n0, r0;\n\t"
        "madc.hi.u32    r1, h0, n0, r1;\n\t"
        "mad.lo.u32     r1, x0, w1, r1;\n\t"
        "mad.lo.u32     r1, x1, w0, r1;\n\t"
        "mad.lo.u32     r1, h0, n1, r1;\n\t"
        "mad.lo.u32     r1, h1, n0, r1;\n\t"
        "mov.b64 %0, {r0, r1};\n\t"
        "}"
        : "=l"(r)
        : "l"(x), "l"(w), "l"(wsh), "l"(nq));
    return r;
}
__device__ __forceinline__ u64 mul_shoup_lazy(u64 x, u64 w, u64 wsh, u64 q) { return mul_shoup_lazy_nq(x, w, wsh, 0ull - q); }
__device__ __forceinline__ u64 mul_shoup(u64 x, u64 w, u64 wsh, u64 q) { return csub(mul_shoup_lazy(x, w, wsh, q), q); }

// x * w mod q with an APPROXIMATE Shoup quotient, lazily in [0, 3q); any x < 2^64.  The quotient drops the lo x lo partial
// product of x * wsh (h' = x1*s1 + floor((x0*s1 + x1*s0) / 2^32) is h or h - 1), which saves one of the four wide
// multiplies of the high product; written as 32-bit carry chains so that ptxas emits 5 wide multiplies + 4 IMAD and the
// additions stay on the ALU pipe (the C form of the exact product compiles to 6 IMAD.WIDE + 4 IMAD + ~1.5 IMAD.X/IADD/MOV
// on the multiplier pipe that bounds the 60-bit transforms).  nq = 2^64 - q.
__device__ __forceinline__ u64 mul_shoup_lazy3_nq(u64 x, u64 w, u64 wsh, u64 nq)
{
    u64 r;
    asm("{\n\t"
        ".reg .u32 x0, x1, w0, w1, s0, s1, n0, n1, m0, m1, m2, h0, h1, r0, r1;\n\t"
        "mov.b64 {x0, x1}, %1;\n\t"
        "mov.b64 {w0, w1}, %2;\n\t"
        "mov.b64 {s0, s1}, %3;\n\t"
        "mov.b64 {n0, n1}, %4;\n\t"
        "mul.lo.u32     m0, x0, s1;\n\t"
        "mul.hi.u32     m1, x0, s1;\n\t"
        "mad.lo.cc.u32  m0, x1, s0, m0;\n\t"
        "madc.hi.cc.u32 m1, x1, s0, m1;\n\t"
        "addc.u32       m2, 0, 0;\n\t"
        "mad.lo.cc.u32  h0, x1, s1, m1;\n\t"
        "madc.hi.u32    h1, x1, s1, m2;\n\t"
        "mul.lo.u32     r0, x0, w0;\n\t"
        "mul.hi.u32     r1, x0, w0;\n\t"
        "mad.lo.cc.u32  r0, h0, n0, r0;\n\t"
        "madc.hi.u32    r1, h0, n0, r1;\n\t"
        "mad.lo.u32     r1, x0, w1, r1;\n\t"
        "mad.lo.u32     r1, x1, w0, r1;\n\t"
        "mad.lo.u32     r1, h0, n1, r1;\n\t"
        "mad.lo.u32     r1, h1, n0, r1;\n\t"
        "mov.b64 %0, {r0, r1};\n\t"
        "}"
        : "=l"(r)
        : "l"(x), "l"(w), "l"(wsh), "l"(nq));
    return r;
}

// x mod q for any x < 2^64 (SEAL barrett_reduce_64)
__device__ __forceinline__ u64 barrett64(u64 x, const ModConst &m)
{
    u64 h = __umul64hi(x, m.mu_hi);
    return csub(x - h * m.q, m.q);
}

// (hi:lo) mod q for hi:lo < 2^124 (SEAL barrett_reduce_128)
__device__ __forceinline__ u64 barrett128(u64 hi, u64 lo, const ModConst &m)
{
    u64 carry = __umul64hi(lo, m.mu_lo);
    u64 t2lo = lo * m.mu_hi, t2hi = __umul64hi(lo, m.mu_hi);
    u64 tmp1 = t2lo + carry;
    u64 tmp3 = t2hi + (tmp1 < t2lo);
    t2lo = hi * m.mu_lo;
    t2hi = __umul64hi(hi, m.mu_lo);
    u64 s = tmp1 + t2lo;
    carry = t2hi + (s < tmp1);
    u64 qhat = hi * m.mu_hi + tmp3 + carry;
    u64 r = lo - qhat * m.q;
    return csub(csub(r, m.q), m.q);
}

__device__ __forceinline__ u64 mulmod(u64 a, u64 b, const ModConst &m)
{
    return barrett128(__umul64hi(a, b), a * b, m);
}

// 128-bit accumulate (hi:lo) += x*y, written as PTX carry chains over the 32-bit halves: ptxas turns each
// mad.lo.cc / madc.hi.cc pair into ONE IMAD.WIDE.U32 with accumulate and carry, so a product costs the four
// multiplies + three IADD3 on the ALU pipe.  (The plain `unsigned __int128` form compiles to the same four
// multiplies plus ~7 add / move instructions, a third of them IMAD.X / IMAD.MOV on the multiplier pipe
// that bounds the multiply-accumulate kernels: measured 9 % slower in dh_inner_kernel.)
// PRECONDITION x, y < 2^63 (residues and lazy values of moduli up to 61 bits are below 2^62): the two cross
// products x0*y1 + x1*y0 then stay below 2^64 and their sum needs no carry of its own.
__device__ __forceinline__ void mac128(u64 &hi, u64 &lo, u64 x, u64 y)
{
    asm("{\n\t"
        ".reg .u32 x0, x1, y0, y1, a0, a1, a2, a3, t0, t1;\n\t"
        "mov.b64 {x0, x1}, %2;\n\t"
        "mov.b64 {y0, y1}, %3;\n\t"
        "mov.b64 {a0, a1}, %0;\n\t"
        "mov.b64 {a2, a3}, %1;\n\t"
        "mad.lo.cc.u32  a0, x0, y0, a0;\n\t"
        "madc.hi.cc.u32 a1, x0, y0, a1;\n\t"
        "madc.lo.cc.u32 a2, x1, y1, a2;\n\t"
        "madc.hi.u32    a3, x1, y1, a3;\n\t"
        "mul.lo.u32     t0, x0, y1;\n\t"
        "mul.hi.u32     t1, x0, y1;\n\t"
        "mad.lo.cc.u32  t0, x1, y0, t0;\n\t"
        "madc.hi.u32    t1, x1, y0, t1;\n\t"
        "add.cc.u32     a1, a1, t0;\n\t"
        "addc.cc.u32    a2, a2, t1;\n\t"
        "addc.u32       a3, a3, 0;\n\t"
        "mov.b64 %0, {a0, a1};\n\t"
        "mov.b64 %1, {a2, a3};\n\t"
        "}"
        : "+l"(lo), "+l"(hi)
        : "l"(x), "l"(y));
}

// Montgomery reduction core: ((hi:lo) + t*q) >> 64 with t = lo * (-q^-1) mod 2^64, congruent to
// (hi:lo) * 2^-64 mod q.  The low 64 bits of the sum cancel; the carry chains below produce the high
// half directly (12 instructions; the mulhi + carry-fix-up form compiles to 17).  For hi:lo < k*q*2^64
// the result is below (k+1)*q.
__device__ __forceinline__ u64 mont_reduce_lazy(u64 hi, u64 lo, const ModConst &m)
{
    const u64 t = lo * m.qinv_neg;
    u64 r;
    asm("{\n\t"
        ".reg .u32 t0, t1, q0, q1, a0, a1, a2, a3, u0, u1, c;\n\t"
        "mov.b64 {t0, t1}, %1;\n\t"
        "mov.b64 {q0, q1}, %2;\n\t"
        "mov.b64 {a0, a1}, %3;\n\t"
        "mov.b64 {a2, a3}, %4;\n\t"
        "mad.lo.cc.u32  a0, t0, q0, a0;\n\t"
        "madc.hi.cc.u32 a1, t0, q0, a1;\n\t"
        "madc.lo.cc.u32 a2, t1, q1, a2;\n\t"
        "madc.hi.u32    a3, t1, q1, a3;\n\t"
        "mul.lo.u32     u0, t0, q1;\n\t"
        "mul.hi.u32     u1, t0, q1;\n\t"
        "mad.lo.cc.u32  u0, t1, q0, u0;\n\t"
        "madc.hi.cc.u32 u1, t1, q0, u1;\n\t"
        "addc.u32       c, 0, 0;\n\t"
        "add.cc.u32     a1, a1, u0;\n\t"
        "addc.cc.u32    a2, a2, u1;\n\t"
        "addc.u32       a3, a3, c;\n\t"
        "mov.b64 %0, {a2, a3};\n\t"
        "}"
        : "=l"(r)
        : "l"(t), "l"(m.q), "l"(lo), "l"(hi));
    return r;
}
// canonical; needs hi:lo < q * 2^64
__device__ __forceinline__ u64 mont_reduce(u64 hi, u64 lo, const ModConst &m) { return csub(mont_reduce_lazy(hi, lo, m), m.q); }
// canonical for sums of up to 32 products of canonical residues (hi:lo < 2 * q * 2^64)
__device__ __forceinline__ u64 mont_reduce_wide(u64 hi, u64 lo, const ModConst &m)
{
    return csub(csub(mont_reduce_lazy(hi, lo, m), m.q), m.q);
}

__device__ __forceinline__ u64 addmod(u64 a, u64 b, u64 q) { return csub(a + b, q); }
__device__ __forceinline__ u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
__device__ __forceinline__ u64 negmod(u64 a, u64 q) { return a ? q - a : 0; }

}  // namespace hegpu
