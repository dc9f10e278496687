// bfly_variants.cu -- static SASS comparison of 60-bit Shoup butterfly formulations (no GPU needed):
//   nvcc -gencode arch=compute_100a,code=sm_100a -cubin -o /tmp/bfly.cubin tools/bfly_variants.cu && cuobjdump -sass /tmp/bfly.cubin
// Each kernel runs 3 radix-2 stages on 8 register-resident coefficients (12 butterflies), like one radix-8 pass.
#include <cstdint>
typedef unsigned long long u64;
typedef unsigned int u32;

__device__ __forceinline__ u64 shoup_c(u64 x, u64 w, u64 wsh, u64 nq)
{
    u64 h = __umul64hi(x, wsh);
    return x * w + h * nq;
}

// approximate quotient: the lo x lo partial product of x * wsh is dropped (h' in {h-1, h}), result in [0, 3q)
__device__ __forceinline__ u64 shoup_ptx(u64 x, u64 w, u64 wsh, u64 nq)
{
    u64 r;
    asm("{\n\t"
        ".reg .u32 x0, x1, w0, w1, s0, s1, n0, n1, m0, m1, m2, h0, h1, r0, r1;\n\t"
        "mov.b64 {x0, x1}, %1;\n\t"
        "mov.b64 {w0, w1}, %2;\n\t"
        "mov.b64 {s0, s1}, %3;\n\t"
        "mov.b64 {n0, n1}, %4;\n\t"
        // mid = x0*s1 + x1*s0 (65 bits: m2:m1:m0)
        "mul.lo.u32     m0, x0, s1;\n\t"
        "mul.hi.u32     m1, x0, s1;\n\t"
        "mad.lo.cc.u32  m0, x1, s0, m0;\n\t"
        "madc.hi.cc.u32 m1, x1, s0, m1;\n\t"
        "addc.u32       m2, 0, 0;\n\t"
        // h = x1*s1 + (m2:m1)
        "mad.lo.cc.u32  h0, x1, s1, m1;\n\t"
        "madc.hi.u32    h1, x1, s1, m2;\n\t"
        // r = lo64(x*w + h*nq)
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

// exact quotient as carry chains (T in [0, 2q))
__device__ __forceinline__ u64 shoup_ptx_exact(u64 x, u64 w, u64 wsh, u64 nq)
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

template <int V>
__device__ __forceinline__ void bfly(u64 &X, u64 &Y, u64 w, u64 wsh, u64 nq, u64 q2)
{
    const u64 T = V == 0 ? shoup_c(Y, w, wsh, nq) : V == 1 ? shoup_ptx(Y, w, wsh, nq) : shoup_ptx_exact(Y, w, wsh, nq);
    Y = X + q2 - T;
    X = X + T;
}

template <int V>
__global__ void pass8(u64 *__restrict__ d, const ulonglong2 *__restrict__ tw, u64 q, u64 q2, int iters)
{
    u64 x[8];
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int k = 0; k < 8; ++k) x[k] = d[t * 8 + k];
    const u64 nq = 0ull - q;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const ulonglong2 W0 = tw[t + it];
        for (int lo = 0; lo < 4; ++lo) bfly<V>(x[lo], x[lo + 4], W0.x, W0.y, nq, q2);
        for (int hi = 0; hi < 2; ++hi) {
            const ulonglong2 W = tw[2 * t + hi + it];
            for (int lo = 0; lo < 2; ++lo) bfly<V>(x[hi * 4 + lo], x[hi * 4 + lo + 2], W.x, W.y, nq, q2);
        }
        for (int hi = 0; hi < 4; ++hi) {
            const ulonglong2 W = tw[4 * t + hi + it];
            bfly<V>(x[hi * 2], x[hi * 2 + 1], W.x, W.y, nq, q2);
        }
        for (int k = 0; k < 8; ++k) x[k] = x[k] >= (q << 3) ? x[k] - (q << 3) : x[k];
    }
    for (int k = 0; k < 8; ++k) d[t * 8 + k] = x[k];
}

template __global__ void pass8<0>(u64 *, const ulonglong2 *, u64, u64, int);
template __global__ void pass8<1>(u64 *, const ulonglong2 *, u64, u64, int);
template __global__ void pass8<2>(u64 *, const ulonglong2 *, u64, u64, int);
