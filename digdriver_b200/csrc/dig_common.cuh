// Shared helpers for libdigb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dig_b200.h"

namespace dig {

void set_error(const char *fmt, ...);
int sm_count();

#define DIG_CHECK_ARG(cond, msg)                           \
    do {                                                   \
        if (!(cond)) {                                     \
            dig::set_error("%s: %s", __func__, msg);       \
            return DIG_ERR_ARG;                            \
        }                                                  \
    } while (0)

#define DIG_CHECK_LAUNCH()                                                          \
    do {                                                                            \
        cudaError_t e__ = cudaGetLastError();                                       \
        if (e__ != cudaSuccess) {                                                   \
            dig::set_error("%s: CUDA error: %s", __func__, cudaGetErrorString(e__)); \
            return DIG_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

#define DIG_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            dig::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__)); \
            return DIG_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

// splitmix64 finaliser -- must stay identical to mix64() in oracle/dig_oracle.c
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

}  // namespace dig
