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
            else if (MODE == 6 || MODE == 7) key = ((x >> (j & 15)) & 2047u) * 4u;       // shared 2048 words random
            else key = ((x >> (j & 15)) & 63u) * 4u;                                     // shared 64 bins random
            const uint32_t addr = base + key;
            if (MODE == 0 || MODE == 2 || MODE == 4)
                asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(1u) : "memory");     // POPC.INC
            else if (MODE == 6)
                asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(((x >> (j + 3)) & 0x10000u) | one) : "memory");   // ADD, data-dependent value
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

// read-and-zero of an 8 KB per-warp region: LDS.128 + STS.128 (MODE 0), ATOMS.EXCH.128 (1), ATOMS.EXCH.64 (2), LDS.128 only (3)
template <int MODE>
__global__ void __launch_bounds__(256) kz(uint32_t *out, int iters)
{
    extern __shared__ __align__(16) uint32_t sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + warp * 8192;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t addr = base + (k * 32 + lane) * 16;
            uint32_t a, b, c, d;
            if (MODE == 0 || MODE == 3) {
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
                if (MODE == 0) asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(addr), "r"(acc & 1u) : "memory");
            } else if (MODE == 1) {
                asm volatile("{.reg .b128 v, z; mov.b128 z, {%5,%5,%5,%5}; atom.shared.exch.b128 v, [%4], z; mov.b128 {%0,%1,%2,%3}, v;}"
                             : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr), "r"(acc & 1u) : "memory");
            } else {
                unsigned long long o0, o1;
                asm volatile("atom.shared.exch.b64 %0, [%1], %2;" : "=l"(o0) : "r"(addr), "l"((unsigned long long)(acc & 1u)) : "memory");
                asm volatile("atom.shared.exch.b64 %0, [%1], %2;" : "=l"(o1) : "r"(addr + 8), "l"((unsigned long long)(acc & 1u)) : "memory");
                a = (uint32_t)o0; b = (uint32_t)(o0 >> 32); c = (uint32_t)o1; d = (uint32_t)(o1 >> 32);
            }
            acc += a + b + c + d;
        }
    }
    if (acc == 0x12345u) out[threadIdx.x] = acc;
}

template <int MODE>
void runz(const char *name, uint32_t *out)
{
    const int iters = 2000, blocks = 148 * 3;
    cudaFuncSetAttribute(kz<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    kz<MODE><<<blocks, 256, 65536>>>(out, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    kz<MODE><<<blocks, 256, 65536>>>(out, iters);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    double passes = (double)blocks * 8 * iters;               // 8 KB passes
    printf("%-44s %8.3f ms  %.1f cycles / 8 KB pass / SM (@1.95GHz)  err=%s\n", name, ms,
           ms * 1e-3 * 1.95e9 / (passes / 148), cudaGetErrorString(cudaGetLastError()));
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
    run<6>("shared 2048 words, ADD 1|bit<<16 (random)", out, keys);
    run<7>("shared 2048 words, ADD reg (random)", out, keys);
    runz<3>("8 KB pass: LDS.128 only", out);
    runz<0>("8 KB pass: LDS.128 + STS.128", out);
    runz<1>("8 KB pass: ATOMS.EXCH.128", out);
    runz<2>("8 KB pass: 2 x ATOMS.EXCH.64", out);
    return 0;
}
