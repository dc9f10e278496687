// mma_micro.cu -- go/no-go for an 8-bit-limb formulation of the double-hoisted inner sums (optional):
// throughput of the legacy warp-level integer MMA (mma.sync.m16n8k32 u8 x u8 -> s32) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_micro tools/mma_micro.cu && ./mma_micro 1965
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned int u32;

template <int CHAINS>
__global__ void __launch_bounds__(256) k(int *out, u32 iters, u32 seed, u32 sink)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    int c[CHAINS][4];
    u32 a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = t * 2654435761u + i + seed;
    b[0] = t ^ seed;
    b[1] = t + seed;
#pragma unroll
    for (int j = 0; j < CHAINS; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) c[j][i] = 0;
    for (u32 it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CHAINS; ++j)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3])
                         : "r"(a[0] + j), "r"(a[1]), "r"(a[2]), "r"(a[3] ^ it), "r"(b[0]), "r"(b[1] + j));
    }
    int acc = 0;
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) acc += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    if (sink) out[t] = acc;
}

template <int CHAINS>
static void run(int sms, double mhz, int *out, int blocks_per_sm)
{
    const u32 iters = 4096, grid = sms * blocks_per_sm, block = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<CHAINS><<<grid, block>>>(out, iters, 12345u + rep, 0u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double mmas = (double)grid * (block / 32) * iters * CHAINS;
    const double macs = mmas * 16 * 8 * 32;
    printf("{\"what\": \"mma.sync.m16n8k32.u8.u8.s32\", \"chains_per_warp\": %d, \"warps_per_sm\": %d, \"mac_per_clk_per_sm\": %.0f, \"TMAC_per_s\": %.1f, "
           "\"cycles_per_mma_per_smsp\": %.2f}\n",
           CHAINS, blocks_per_sm * 8, macs / (best * 1e-3) / (mhz * 1e6) / sms, macs / (best * 1e-3) / 1e12,
           (best * 1e-3) * (mhz * 1e6) / (mmas / sms / 4));
}

int main(int argc, char **argv)
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const double mhz = argc > 1 ? atof(argv[1]) : 1965.0;
    int *out;
    cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    run<1>(p.multiProcessorCount, mhz, out, 2);
    run<4>(p.multiProcessorCount, mhz, out, 2);
    run<8>(p.multiProcessorCount, mhz, out, 2);
    run<8>(p.multiProcessorCount, mhz, out, 4);
    run<15>(p.multiProcessorCount, mhz, out, 2);
    return 0;
}
