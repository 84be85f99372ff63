// K1: ASCII genome -> 2-bit packed bases (MSB first) + N bitmask.
//
// HBM-bound streaming kernel: 1 B/base read, 0.375 B/base written.  Each thread converts
// 16 bases (one 128-bit load) into one packed word with byte-SIMD (__vcmpeq4) and a pair of
// lanes assembles one N-mask word through a shuffle, so loads and both stores are coalesced.
#include "dig_common.cuh"

namespace {

// 4 ASCII bytes (little endian: lowest byte = first base) -> 8 bits of 2-bit codes, first base
// most significant, and 4 validity bits in the same order (1 = not ACGT).
//
// The kernel is ALU-bound (the byte compares of the obvious formulation are emulated), so everything is done on
// whole 32-bit words:
//  * validity: upper-cased letters are 0x41 0x43 0x47 0x54, i.e. (c >> 3) must equal LUT[c & 7] with
//    LUT = {1: 8, 3: 8, 7: 8, 4: 10}; the 8-entry byte LUT is ONE PRMT for all four bytes, a zero-byte test of
//    ((c >> 3) ^ LUT[c & 7]) gives the four validity bits;
//  * codes: ((c >> 1) & 3) ^ ((c >> 2) & 1) -> 0,1,2,3; the four 2-bit fields and the four validity bits are gathered
//    with one multiply each (no partial product of the magic constants carries into the result byte);
//  * non-ACGT letters other than N are counted only in words that contain an invalid byte (N runs are rare).
__device__ __forceinline__ void convert4(uint32_t w, uint32_t &codes8, uint32_t &n4, uint32_t &other)
{
    const uint32_t up = w & 0xDFDFDFDFu;                       // upper-case
    // selector nibbles (c & 7) of the four bytes packed into 16 bits
    const uint32_t lo3 = w & 0x07070707u;
    const uint32_t y = lo3 | (lo3 >> 4);
    const uint32_t sel = (y & 0xFFu) | ((y >> 8) & 0xFF00u);
    // LUT bytes 0..7 = FF 08 FF 08 0A FF FF 08 (index = c & 7)
    const uint32_t want = __byte_perm(0x08FF08FFu, 0x08FFFF0Au, sel);
    const uint32_t d = ((up >> 3) & 0x1F1F1F1Fu) ^ want;       // byte == 0 <=> valid letter
    const uint32_t nz = ((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d; // bit 7 of each byte: byte != 0
    const uint32_t bad = (nz >> 7) & 0x01010101u;              // 1 per invalid byte
    uint32_t t = ((w >> 1) & 0x03030303u) ^ ((w >> 2) & 0x01010101u);
    t &= (bad ^ 0x01010101u) * 3u;                             // non-ACGT stored as 0
    codes8 = (t * 0x40100401u) >> 24;                          // t0<<6 | t1<<4 | t2<<2 | t3
    n4 = (bad * 0x08040201u) >> 24;                            // b0<<3 | b1<<2 | b2<<1 | b3
    if (bad) {
        const uint32_t x = up ^ 0x4E4E4E4Eu;                   // byte == 0 <=> 'N' / 'n'
        const uint32_t notn = ((((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) >> 7) & 0x01010101u;
        other += __popc(bad & notn);
    }
}

constexpr int PACK_UNROLL = 4;     // independent 128-bit loads in flight per thread

__device__ __forceinline__ void load_group(const uint8_t *__restrict__ ascii, int64_t n, int64_t gidx, int aligned,
                                           uint32_t (&w)[4])
{
    const int64_t b0 = gidx << 4;
    if (aligned && b0 + 16 <= n) {
        const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(ascii + b0));      // streaming: read once
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t x = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int64_t g = b0 + q * 4 + b;
                const uint32_t c = g < n ? ascii[g] : (uint32_t)'N';
                x |= c << (8 * b);
            }
            w[q] = x;
        }
    }
}

__global__ void __launch_bounds__(256) pack_kernel(const uint8_t *__restrict__ ascii, int64_t n,
                                                   uint32_t *__restrict__ packed2, uint32_t *__restrict__ nmask,
                                                   unsigned long long *n_other, int64_t n_groups, int aligned)
{
    // n_groups is even; group gidx covers bases [16*gidx, 16*gidx+16).  A warp takes PACK_UNROLL runs of 32
    // consecutive groups per iteration; the loop condition is warp-uniform so the pair shuffle sees a full warp.
    const int lane = threadIdx.x & 31;
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    uint32_t other = 0;
    for (int64_t base = warp_id * (32 * PACK_UNROLL); base < n_groups; base += n_warps * (32 * PACK_UNROLL)) {
        uint32_t w[PACK_UNROLL][4];
#pragma unroll
        for (int j = 0; j < PACK_UNROLL; ++j) load_group(ascii, n, base + j * 32 + lane, aligned, w[j]);
#pragma unroll
        for (int j = 0; j < PACK_UNROLL; ++j) {
            const int64_t gidx = base + j * 32 + lane;
            const bool live = gidx < n_groups;
            uint32_t word = 0, n16 = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t c8, n4;
                convert4(w[j][q], c8, n4, other);
                word |= c8 << (24 - 8 * q);
                n16 |= n4 << (12 - 4 * q);
            }
            // positions beyond n were fed as 'N': they are flagged in the mask and do not count as "other"
            if (live) __stcs(packed2 + gidx, word);
            const uint32_t peer = __shfl_xor_sync(0xffffffffu, n16, 1);
            if (live && (gidx & 1) == 0) __stcs(nmask + (gidx >> 1), (n16 << 16) | peer);
        }
    }
    if (n_other) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) other += __shfl_xor_sync(0xffffffffu, other, o);
        if (lane == 0 && other) atomicAdd(n_other, (unsigned long long)other);
    }
}

}  // namespace

extern "C" {

int64_t dig_packed_words(int64_t n_bases) { return ((n_bases + 31) >> 5) << 1; }
int64_t dig_nmask_words(int64_t n_bases) { return (n_bases + 31) >> 5; }

int dig_pack_genome(const uint8_t *ascii_d, int64_t n_bases, uint32_t *packed2_d, uint32_t *nmask_d,
                    unsigned long long *n_other_d, void *stream)
{
    DIG_CHECK_ARG(n_bases >= 0, "negative size");
    if (n_bases == 0) return DIG_OK;
    DIG_CHECK_ARG(ascii_d && packed2_d && nmask_d, "null pointer");
    const int64_t n_groups = ((n_bases + 31) >> 5) << 1;      // two 16-base groups per N-mask word
    const int aligned = (reinterpret_cast<uintptr_t>(ascii_d) & 15u) == 0;
    const int threads = 256;
    int64_t blocks = (n_groups + threads * PACK_UNROLL - 1) / (threads * PACK_UNROLL);
    const int64_t cap = (int64_t)dig::sm_count() * 8;
    if (blocks > cap) blocks = cap;
    pack_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(ascii_d, n_bases, packed2_d, nmask_d,
                                                                        n_other_d, n_groups, aligned);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

}
