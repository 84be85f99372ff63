// Micro-benchmark: FP64 pipe peak of this GPU (DFMA, and the DMUL/DADD mix the NB tail sums are made of), the
// denominator for the `fp64_frac` figures of the test-stage kernels (K7 gene test, per-site test, K8 position test).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_dfma micro_dfma.cu && ./micro_dfma
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double c1, double c2)
{
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] = fma(a[i], c1, c2);                       // DFMA, 8 independent chains
                else if (MODE == 1) a[i] = a[i] * c1;                          // DMUL
                else if (MODE == 2) a[i] = a[i] + c2;                          // DADD
                else a[i] = (i & 1) ? fma(a[i], c1, c2) : a[i] * c1 + a[(i + 1) & 7];   // mixed, cross-chain
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 0.12345) out[threadIdx.x] = s;
}

template <int MODE>
double run(const char *name, double *out, int sms, double flop_per_instr)
{
    const int iters = 2000, blocks = sms * 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<blocks, 256>>>(out, 10, 0.999999, 1e-7);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        k<MODE><<<blocks, 256>>>(out, iters, 0.999999, 1e-7);
        cudaEventRecord(b);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double instr = (double)blocks * 256 * iters * 64;        // thread instructions
    const double rate = instr / (best * 1e-3);
    printf("%-24s %8.3f ms  %8.2f G thread-instr/s  %7.2f TFLOP/s  %.2f instr / cycle / SM at 1.965 GHz\n", name, best,
           rate / 1e9, rate * flop_per_instr / 1e12, rate / sms / 1.965e9);
    return rate;
}

int main()
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, dev);
    printf("%s, %d SMs\n", p.name, sms);
    double *out;
    cudaMalloc(&out, 4096);
    const double dfma = run<0>("DFMA", out, sms, 2.0);
    run<1>("DMUL", out, sms, 1.0);
    run<2>("DADD", out, sms, 1.0);
    run<3>("DFMA/DMUL+DADD mix", out, sms, 1.5);
    printf("fp64_peak_dfma_per_s %.6e\n", dfma);
    return 0;
}
