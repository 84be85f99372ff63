// Synthetic genome generator: the device twin of orc_synth_genome (oracle/dig_oracle.c).
// Position g is a pure function of (seed, g), so any slice can be produced on any rank.
#include "dig_common.cuh"

namespace {

__device__ __forceinline__ uint8_t synth_base(uint64_t g, uint64_t seed, int n_frac16)
{
    const uint64_t h = dig::mix64(seed ^ (g * 0xD1342543DE82EF95ull));
    uint32_t ch = (0x54474341u >> (8 * (uint32_t)(h & 3))) & 0xFFu;   // "ACGT"[h & 3]
    if ((h >> 2) & 1) ch |= 0x20u;
    const uint64_t sb = g >> 20;
    const uint64_t hs = dig::mix64(seed ^ 0xA5A5A5A5ull ^ (sb * 0x9E3779B97F4A7C15ull));
    const uint64_t off = hs % (uint64_t)((1 << 20) - 51200);
    const uint64_t len = ((10240 + (hs >> 32) % 40960) * (uint64_t)n_frac16) >> 4;
    const uint64_t q = g & ((1u << 20) - 1);
    if (q >= off && q < off + len) ch = 'N';
    return (uint8_t)ch;
}

__global__ void __launch_bounds__(256) synth_kernel(uint8_t *__restrict__ out, int64_t g0, int64_t n,
                                                    uint64_t seed, int n_frac16)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 16;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; i < n; i += stride) {
        if (i + 16 <= n && ((reinterpret_cast<uintptr_t>(out + i) & 15u) == 0)) {
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t x = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    x |= (uint32_t)synth_base((uint64_t)(g0 + i + q * 4 + b), seed, n_frac16) << (8 * b);
                w[q] = x;
            }
            *reinterpret_cast<uint4 *>(out + i) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            for (int b = 0; b < 16 && i + b < n; ++b)
                out[i + b] = synth_base((uint64_t)(g0 + i + b), seed, n_frac16);
        }
    }
}

}  // namespace

extern "C" int dig_synth_genome(uint8_t *ascii_d, int64_t g0, int64_t n, uint64_t seed, int n_frac16, void *stream)
{
    DIG_CHECK_ARG(n >= 0 && g0 >= 0, "negative size/offset");
    if (n == 0) return DIG_OK;
    DIG_CHECK_ARG(ascii_d != nullptr, "null pointer");
    const int threads = 256;
    int64_t blocks = (n / 16 + threads) / threads;
    const int64_t cap = (int64_t)dig::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    synth_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(ascii_d, g0, n, seed, n_frac16);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}
