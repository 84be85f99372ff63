// Micro-benchmark: issue rate of the integer instructions the scan kernels are made of (B200).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_pipes micro_pipes.cu && ./micro_pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(uint32_t *out, int iters, uint32_t c1, uint32_t c2)
{
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 2654435761u + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] = __funnelshift_r(a[i], a[(i + 1) & 7], 7);                // SHF
                else if (MODE == 1) asm volatile("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(a[i]) : "r"(a[i]), "r"(c1), "r"(a[(i + 1) & 7]));   // LOP3
                else if (MODE == 2) asm volatile("prmt.b32.f4e %0, %1, %2, %3;" : "=r"(a[i]) : "r"(a[i]), "r"(c1), "r"(a[(i + 1) & 7]));
                else if (MODE == 3) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a[i]) : "r"(a[i]), "r"(c1), "r"(a[(i + 1) & 7]));   // IMAD
                else if (MODE == 4) asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(a[i]) : "r"(a[i]), "r"(c2));                             // IMAD.HI
                else if (MODE == 5) a[i] = __dp4a(a[i], c1, a[(i + 1) & 7]);                   // IDP.4A
                else if (MODE == 6) asm volatile("add.u32 %0, %1, %2;" : "=r"(a[i]) : "r"(a[i]), "r"(a[(i + 1) & 7]));                     // IADD
                else if (MODE == 7) {                                                          // SHF + IMAD interleaved
                    if (i & 1) a[i] = __funnelshift_r(a[i], a[(i + 1) & 7], 7);
                    else asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a[i]) : "r"(a[i]), "r"(c1), "r"(a[(i + 1) & 7]));
                } else if (MODE == 8) {                                                        // SHF + IMAD.HI interleaved
                    if (i & 1) a[i] = __funnelshift_r(a[i], a[(i + 1) & 7], 7);
                    else asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(a[i]) : "r"(a[i]), "r"(c2));
                } else if (MODE == 9) {                                                        // LOP3 + IDP interleaved
                    if (i & 1) asm volatile("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(a[i]) : "r"(a[i]), "r"(c1), "r"(a[(i + 1) & 7]));
                    else a[i] = __dp4a(a[i], c1, a[(i + 1) & 7]);
                }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    if (s == 0x12345u) out[threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, uint32_t *out)
{
    const int iters = 4000, blocks = 148 * 2;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<blocks, 512>>>(out, 10, 32u, 4096u);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<MODE><<<blocks, 512>>>(out, iters, 32u, 4096u);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double winstr = (double)blocks * 16 * iters * 64;        // warp instructions
    const double cyc = ms * 1e-3 * 1.95e9;
    printf("%-28s %8.3f ms  %.2f warp-instr / cycle / SM  (%.2f per SMSP)\n", name, ms, winstr / 148 / cyc, winstr / 148 / cyc / 4);
}

int main()
{
    uint32_t *out;
    cudaMalloc(&out, 4096);
    run<0>("SHF", out);
    run<1>("LOP3", out);
    run<2>("PRMT.F4E", out);
    run<3>("IMAD (const operand)", out);
    run<4>("IMAD.HI", out);
    run<5>("IDP.4A", out);
    run<6>("IADD3", out);
    run<7>("SHF + IMAD", out);
    run<8>("SHF + IMAD.HI", out);
    run<9>("LOP3 + IDP.4A", out);
    return 0;
}
