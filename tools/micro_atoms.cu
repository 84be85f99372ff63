// Micro-benchmark: shared-memory atomic throughput on B200 for the access patterns the scan kernel can use.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_atoms micro_atoms.cu && ./micro_atoms
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, const uint32_t *keys, int iters, uint32_t one)
{
    extern __shared__ uint32_t sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + warp * 8192;   // 8 KB per warp
    uint32_t x = keys[blockIdx.x * blockDim.x + threadIdx.x];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            uint32_t key;
            if (MODE == 0 || MODE == 1 || MODE == 5) key = ((x >> (j & 15)) & 63u) * 128u + lane * 4u;   // private, conflict free
            else if (MODE == 2 || MODE == 3) key = ((x >> (j & 15)) & 1023u) * 4u;       // shared 1024 bins random
            else key = ((x >> (j & 15)) & 63u) * 4u;                                     // shared 64 bins random
            const uint32_t addr = base + key;
            if (MODE == 0 || MODE == 2 || MODE == 4)
                asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(1u) : "memory");     // POPC.INC
            else if (MODE == 5)
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(x) : "memory");          // plain store
            else
                asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(one) : "memory");    // ADD reg
        }
        x = x * 1664525u + 1013904223u;
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = sm[5] + sm[4097];
}

template <int MODE>
void run(const char *name, uint32_t *out, uint32_t *keys)
{
    const int iters = 2000, blocks = 148 * 3;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<blocks, 256, 65536>>>(out, keys, 10, 1u);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<MODE><<<blocks, 256, 65536>>>(out, keys, iters, 1u);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    double instr = (double)blocks * 8 * iters * 32;           // warp-level instructions
    double per_sm_cyc = ms * 1e-3 * 1.95e9;                   // approx cycles at 1.95 GHz
    printf("%-44s %8.3f ms  %.2f cycles/warp-instr/SM (@1.95GHz)  %.1f Glane-ops/s  err=%s\n", name, ms,
           per_sm_cyc / (instr / 148), instr * 32 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    uint32_t *out, *keys;
    cudaMalloc(&out, 4096 * 4);
    cudaMalloc(&keys, 148 * 3 * 256 * 4);
    uint32_t *h = new uint32_t[148 * 3 * 256];
    uint32_t s = 12345;
    for (int i = 0; i < 148 * 3 * 256; ++i) { s = s * 1103515245u + 12345u; h[i] = s ^ (s >> 13); }
    cudaMemcpy(keys, h, 148 * 3 * 256 * 4, cudaMemcpyHostToDevice);
    run<0>("private 64 bins, POPC.INC (conflict-free)", out, keys);
    run<1>("private 64 bins, ADD reg   (conflict-free)", out, keys);
    run<5>("private layout, plain STS  (conflict-free)", out, keys);
    run<2>("shared 1024 bins, POPC.INC (random)", out, keys);
    run<3>("shared 1024 bins, ADD reg   (random)", out, keys);
    run<4>("shared 64 bins, POPC.INC   (random)", out, keys);
    return 0;
}
