// pipe_micro.cu -- instruction-throughput microbenchmarks behind DESIGN.md's pipe model (measurement tool, not product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_micro tools/pipe_micro.cu && ./pipe_micro
// Each kernel runs 8 independent dependent-chains per thread of one instruction (or a fixed mix), operands that change
// every iteration (nothing loop-invariant to hoist), 8 CTAs x 256 threads per SM.  Output: thread-level operations per
// clock per SM and warp-level issue cycles per operation per SM sub-partition.  Check the loop bodies with
// `cuobjdump -sass pipe_micro` -- ptxas is free to pick other opcodes than the PTX suggests.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

#define CHAINS 8
#define REPS 4

template <int KIND>
__global__ void __launch_bounds__(256) k(u64 *out, u32 iters, u32 seed, u32 sink)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    u32 r[CHAINS], s[CHAINS];
    u64 a[CHAINS];
    double d[CHAINS];
    const u32 y = (t ^ seed) | 1u;
    const double dy = 1.0 + 1e-9 * (double)(y & 1023u), dz = 1e-9 * (double)(seed & 1023u);
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        r[i] = t * 2654435761u + i + seed;
        s[i] = t + 77u * i + seed;
        a[i] = ((u64)r[i] << 32) | s[i];
        d[i] = (double)i + 0.5;
    }
    for (u32 it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < REPS; ++rep) {
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) {
                const int j = (i + 1) % CHAINS;
                if (KIND == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(r[j]), "r"(y));                 // IMAD
                if (KIND == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(r[j]), "r"(y));                 // IMAD.HI.U32
                if (KIND == 2) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(a[i]) : "r"((u32)(a[i] >> 32)), "r"((u32)a[j]));  // IMAD.WIDE.U32 .. RZ, both halves live
                if (KIND == 3) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[i]) : "r"((u32)a[j]), "r"((u32)(a[j] >> 32)));  // IMAD.WIDE.U32 with 64-bit addend
                if (KIND == 4) asm volatile("add.u32 %0, %0, %1;\n\txor.b32 %0, %0, %2;" : "+r"(r[i]) : "r"(r[j]), "r"(y)); // 2 ALU ops
                if (KIND == 5) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dy), "d"(dz));                  // DFMA
                if (KIND == 6) {  // IMAD.WIDE (no addend) + 2 independent ALU ops
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(a[i]) : "r"((u32)(a[i] >> 32)), "r"((u32)a[j]));
                    asm volatile("add.u32 %0, %0, %1;\n\txor.b32 %0, %0, %2;" : "+r"(s[i]) : "r"(s[j]), "r"(y));
                }
                if (KIND == 7) {  // IMAD + 2 independent ALU ops
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(r[j]), "r"(y));
                    asm volatile("add.u32 %0, %0, %1;\n\txor.b32 %0, %0, %2;" : "+r"(s[i]) : "r"(s[j]), "r"(y));
                }
                if (KIND == 8) {  // DFMA + IMAD (do the FP64 and the multiplier pipe overlap?)
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dy), "d"(dz));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(r[j]), "r"(y));
                }
                if (KIND == 9) {  // DFMA + IMAD.WIDE
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dy), "d"(dz));
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(a[i]) : "r"((u32)(a[i] >> 32)), "r"((u32)a[j]));
                }
                if (KIND == 10) {  // mul.lo + mul.hi of the same operands (the 64-bit product as two 32-bit instructions)
                    u32 lo, hi;
                    asm volatile("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(r[i]), "r"(s[j]));
                    r[i] = lo ^ y;
                    s[i] = hi + y;
                }
                if (KIND == 11) {  // carry-fused pair: mad.lo.cc + madc.hi (what mac128 is built from)
                    u32 lo = (u32)a[i], hi = (u32)(a[i] >> 32);
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"((u32)a[j]), "r"((u32)(a[j] >> 32)));
                    a[i] = ((u64)hi << 32) | lo;
                }
                if (KIND == 12) {  // DFMA + 2 ALU ops
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dy), "d"(dz));
                    asm volatile("add.u32 %0, %0, %1;\n\txor.b32 %0, %0, %2;" : "+r"(s[i]) : "r"(s[j]), "r"(y));
                }
            }
        }
    }
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc += a[i] + r[i] + s[i] + (u64)__double_as_longlong(d[i]);
    if (sink) out[t] = acc;
}

template <int KIND>
static void run(const char *name, int sms, double mhz, u64 *out)
{
    const u32 iters = 1024, grid = sms * 8, block = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<KIND><<<grid, block>>>(out, iters, 12345u + rep, 0u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double ops = (double)grid * block * iters * REPS * CHAINS;
    const double per_clk_sm = ops / (best * 1e-3) / (mhz * 1e6) / sms;
    printf("{\"kind\": %d, \"what\": \"%s\", \"ops_per_s\": %.4g, \"ops_per_clk_per_sm\": %.2f, \"issue_cycles_per_warp_op_per_smsp\": %.2f}\n", KIND, name,
           ops / (best * 1e-3), per_clk_sm, 32.0 / (per_clk_sm / 4.0));
}

int main(int argc, char **argv)
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = argc > 1 ? atof(argv[1]) : khz / 1000.0;
    u64 *out;
    cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 8);
    printf("{\"device\": \"%s\", \"sms\": %d, \"sm_mhz_assumed\": %.0f}\n", p.name, p.multiProcessorCount, mhz);
    const int sms = p.multiProcessorCount;
    run<0>("IMAD (mad.lo.u32)", sms, mhz, out);
    run<1>("IMAD.HI.U32 (mad.hi.u32)", sms, mhz, out);
    run<2>("IMAD.WIDE.U32 no addend (mul.wide.u32)", sms, mhz, out);
    run<3>("mad.wide.u32 with 64-bit addend", sms, mhz, out);
    run<4>("2 ALU ops (add + xor)", sms, mhz, out);
    run<5>("DFMA", sms, mhz, out);
    run<6>("IMAD.WIDE + 2 ALU", sms, mhz, out);
    run<7>("IMAD + 2 ALU", sms, mhz, out);
    run<8>("DFMA + IMAD", sms, mhz, out);
    run<9>("DFMA + IMAD.WIDE", sms, mhz, out);
    run<10>("mul.lo + mul.hi pair", sms, mhz, out);
    run<11>("mad.lo.cc + madc.hi pair", sms, mhz, out);
    run<12>("DFMA + 2 ALU", sms, mhz, out);
    return 0;
}
