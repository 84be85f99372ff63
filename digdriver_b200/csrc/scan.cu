// K2 / K4: per-region k-mer context histograms over the 2-bit packed genome.
//
// Design (B200): one WARP owns one region (a 10 kb window is 313 32-base words, i.e. ten
// fully coalesced 384-byte warp loads) and a private K-bin int32 histogram in shared memory,
// so there is no cross-warp contention and only __syncwarp() is ever needed.  Every lane
// takes one 32-base word per iteration (one 64-bit load of bases + one 32-bit load of the
// N mask), gets the 2-base halo of its neighbours by shuffle, and issues 32 shared-memory
// atomics whose k-mer indices are produced by one funnel shift + one mask each (all shift
// amounts are compile-time constants).  Validity (window range + "k-mer contains N") is a
// 32-bit mask built with a handful of shifts; when the whole warp is fully valid the atomics
// run unpredicated.  The histogram is written out with 128-bit coalesced stores, re-zeroed
// in the same pass, and folded into per-lane register totals that become the genome-wide
// context totals (DigPreprocess.py:59) without re-reading the counts.
//
// HBM traffic per base: 0.25 B bases + 0.125 B mask + 4K/W B counts.
#include "dig_common.cuh"

namespace {

constexpr int WARPS_PER_BLOCK = 8;
constexpr int THREADS = WARPS_PER_BLOCK * 32;

// reverse-complement of a klen-base k-mer index (5' base most significant)
__device__ __forceinline__ uint32_t revcomp_key(uint32_t key, int klen)
{
    uint32_t x = __brev(key);                                    // reverses pairs AND bits inside pairs
    x = ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);     // undo the swap inside each pair
    x >>= (32 - 2 * klen);
    return x ^ ((1u << (2 * klen)) - 1u);                        // complement: b -> 3 - b
}

struct RegionSpan {
    int64_t gs, ge;   // global centre range [gs, ge)
};

// Centre range of region r exactly as the reference walks it (see dig_b200.h); u/d are the
// number of bases taken to the left/right of the centre on the PLUS strand.
__device__ __forceinline__ RegionSpan region_span(const int64_t *__restrict__ chrom_off,
                                                  const int64_t *__restrict__ chrom_len,
                                                  const int32_t *__restrict__ reg_chrom,
                                                  const int64_t *__restrict__ reg_start,
                                                  const int64_t *__restrict__ reg_end, int64_t r, int n_up,
                                                  int n_down, int u, int d)
{
    const int32_t c = __ldg(reg_chrom + r);
    const int64_t L = __ldg(chrom_len + c);
    const int64_t off = __ldg(chrom_off + c);
    int64_t s = __ldg(reg_start + r);
    const int64_t e = __ldg(reg_end + r);
    if (s < n_up) s = n_up;                 // START == 0 -> n_up (sequence_tools.py:25-26)
    int64_t f0 = s - n_up;                  // fetched string [f0, f1)
    int64_t f1 = e + n_down;
    if (f1 > L) f1 = L;                     // faidx clips at the chromosome end
    if (f0 > L) f0 = L;
    RegionSpan sp;
    sp.gs = off + f0 + u;
    sp.ge = off + f1 - d;
    if (sp.ge < sp.gs) sp.ge = sp.gs;
    return sp;
}

template <int U, int D, int I>
__device__ __forceinline__ uint32_t extract_key(uint32_t a, uint32_t b0, uint32_t b1, uint32_t c)
{
    constexpr int KLEN = U + D + 1;
    constexpr uint32_t MASK = (1u << (2 * KLEN)) - 1u;
    constexpr int BO = 2 * (16 + I - U);        // bit offset from the MSB of [a|b0|b1|c]
    constexpr int Q = BO >> 5;
    constexpr int R = BO & 31;
    const uint32_t hi = Q == 0 ? a : (Q == 1 ? b0 : b1);
    const uint32_t lo = Q == 0 ? b0 : (Q == 1 ? b1 : c);
    if constexpr (R + 2 * KLEN <= 32) {
        return (hi >> (32 - R - 2 * KLEN)) & MASK;
    } else {
        return __funnelshift_r(lo, hi, 64 - R - 2 * KLEN) & MASK;
    }
}

template <int U, int D, int I>
struct Unroll {
    template <bool PRED>
    static __device__ __forceinline__ void run(int *hist, uint32_t a, uint32_t b0, uint32_t b1, uint32_t c,
                                               uint32_t valid)
    {
        if (!PRED || (valid & (0x80000000u >> I))) atomicAdd(hist + extract_key<U, D, I>(a, b0, b1, c), 1);
        Unroll<U, D, I + 1>::template run<PRED>(hist, a, b0, b1, c, valid);
    }
};
template <int U, int D>
struct Unroll<U, D, 32> {
    template <bool PRED>
    static __device__ __forceinline__ void run(int *, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t) {}
};

// ---------------------------------------------------------------------------------------
// fast kernel: symmetric context (U == D), K = 4^(2U+1) <= 1024
// ---------------------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(THREADS) scan_sym_kernel(
    const uint2 *__restrict__ p2v, const uint32_t *__restrict__ p2, const uint32_t *__restrict__ nmask,
    int64_t n_words32, const int64_t *__restrict__ chrom_off, const int64_t *__restrict__ chrom_len,
    const int32_t *__restrict__ reg_chrom, const int64_t *__restrict__ reg_start,
    const int64_t *__restrict__ reg_end, const int8_t *__restrict__ reg_strand, int64_t n_reg,
    int32_t *__restrict__ counts, unsigned long long *__restrict__ totals)
{
    constexpr int D = U;
    constexpr int KLEN = 2 * U + 1;
    constexpr int K = 1 << (2 * KLEN);
    constexpr int K4 = K / 4;                         // int4 chunks (K >= 4)
    constexpr int NTOT = (K4 + 31) / 32;              // int4 chunks per lane
    extern __shared__ __align__(16) int smem[];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    int *hist = smem + warp * K;
    for (int k = lane; k < K; k += 32) hist[k] = 0;
    __syncwarp();

    int tot[NTOT * 4];
#pragma unroll
    for (int j = 0; j < NTOT * 4; ++j) tot[j] = 0;
    int64_t tot_bases = 0;

    const int64_t gwarp = (int64_t)blockIdx.x * WARPS_PER_BLOCK + warp;
    const int64_t nwarps = (int64_t)gridDim.x * WARPS_PER_BLOCK;

    for (int64_t r = gwarp; r < n_reg; r += nwarps) {
        const RegionSpan sp = region_span(chrom_off, chrom_len, reg_chrom, reg_start, reg_end, r, U, D, U, D);
        const bool minus = reg_strand != nullptr && __ldg(reg_strand + r) < 0;
        if (sp.ge > sp.gs) {
            const int64_t w0 = sp.gs >> 5;
            const int64_t w1 = (sp.ge - 1) >> 5;
            for (int64_t wb = w0; wb <= w1; wb += 32) {
                const int64_t w = wb + lane;
                const bool loadable = w < n_words32;
                uint2 pw = make_uint2(0u, 0u);
                uint32_t nm = 0xFFFFFFFFu;
                if (loadable) {
                    pw = __ldg(p2v + w);
                    nm = __ldg(nmask + w);
                }
                uint32_t prev_p = __shfl_up_sync(0xffffffffu, pw.y, 1);
                uint32_t next_p = __shfl_down_sync(0xffffffffu, pw.x, 1);
                uint32_t prev_n = __shfl_up_sync(0xffffffffu, nm, 1);
                uint32_t next_n = __shfl_down_sync(0xffffffffu, nm, 1);
                if (lane == 0) {
                    prev_p = w > 0 ? __ldg(p2 + 2 * w - 1) : 0u;
                    prev_n = w > 0 ? __ldg(nmask + w - 1) : 0xFFFFFFFFu;
                }
                if (lane == 31) {
                    const bool ok = w + 1 < n_words32;
                    next_p = ok ? __ldg(p2 + 2 * w + 2) : 0u;
                    next_n = ok ? __ldg(nmask + w + 1) : 0xFFFFFFFFu;
                }
                // positions of this word that are centres of the region
                const int64_t base_g = w << 5;
                const int64_t lo64 = sp.gs - base_g;
                const int64_t hi64 = sp.ge - base_g;
                const int lo = lo64 < 0 ? 0 : (lo64 > 32 ? 32 : (int)lo64);
                const int hi = hi64 < 0 ? 0 : (hi64 > 32 ? 32 : (int)hi64);
                uint32_t valid = 0u;
                if (hi > lo) {
                    const uint32_t from_lo = lo >= 32 ? 0u : (0xFFFFFFFFu >> lo);
                    const uint32_t below_hi = hi >= 32 ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> hi);
                    valid = from_lo & below_hi;
                }
                // k-mers touching a non-ACGT base are skipped
                uint32_t bad = nm;
#pragma unroll
                for (int t = 1; t <= D; ++t) bad |= (nm << t) | (next_n >> (32 - t));
#pragma unroll
                for (int t = 1; t <= U; ++t) bad |= (nm >> t) | (prev_n << (32 - t));
                valid &= ~bad;

                if (__all_sync(0xffffffffu, valid == 0xFFFFFFFFu)) {
                    Unroll<U, D, 0>::template run<false>(hist, prev_p, pw.x, pw.y, next_p, valid);
                } else if (__any_sync(0xffffffffu, valid != 0u)) {
                    Unroll<U, D, 0>::template run<true>(hist, prev_p, pw.x, pw.y, next_p, valid);
                }
            }
        }
        __syncwarp();
        // write-out: counts row r, re-zero the histogram, fold into the register totals
        int32_t *out = counts + r * (int64_t)K;
        if (!minus) {
            int4 *hist4 = reinterpret_cast<int4 *>(hist);
            int4 *out4 = reinterpret_cast<int4 *>(out);
#pragma unroll
            for (int j = 0; j < NTOT; ++j) {
                const int cidx = j * 32 + lane;
                if (K4 >= 32 || cidx < K4) {
                    const int4 v = hist4[cidx];
                    hist4[cidx] = make_int4(0, 0, 0, 0);
                    __stcs(out4 + cidx, v);
                    tot[4 * j + 0] += v.x;
                    tot[4 * j + 1] += v.y;
                    tot[4 * j + 2] += v.z;
                    tot[4 * j + 3] += v.w;
                }
            }
        } else {
            // minus strand: the reverse-complemented string has the reverse-complemented k-mers
            for (int k = lane; k < K; k += 32) {
                const int v = hist[k];
                hist[k] = 0;
                const uint32_t rk = revcomp_key((uint32_t)k, KLEN);
                out[rk] = v;
                if (totals != nullptr && v) atomicAdd(totals + rk, (unsigned long long)v);
            }
        }
        __syncwarp();
        tot_bases += sp.ge - sp.gs;
        if (tot_bases > (int64_t)1 << 30) {          // keep the int32 register totals from overflowing
            if (totals != nullptr) {
#pragma unroll
                for (int j = 0; j < NTOT; ++j) {
                    const int cidx = j * 32 + lane;
                    if (K4 >= 32 || cidx < K4) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (tot[4 * j + q]) atomicAdd(totals + 4 * cidx + q, (unsigned long long)tot[4 * j + q]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < NTOT * 4; ++j) tot[j] = 0;
            tot_bases = 0;
        }
    }

    if (totals == nullptr) return;                  // uniform across the grid
    // block-level reduction of the totals in shared memory (aliases the now all-zero histograms)
    __syncthreads();
    unsigned long long *ctot = reinterpret_cast<unsigned long long *>(smem);
    static_assert(WARPS_PER_BLOCK * 4 >= 8, "ctot must fit in the histogram area");
    for (int k = threadIdx.x; k < K; k += THREADS) ctot[k] = 0ull;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NTOT; ++j) {
        const int cidx = j * 32 + lane;
        if (K4 >= 32 || cidx < K4) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (tot[4 * j + q]) atomicAdd(ctot + 4 * cidx + q, (unsigned long long)tot[4 * j + q]);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += THREADS)
        if (ctot[k]) atomicAdd(totals + k, ctot[k]);
}

// ---------------------------------------------------------------------------------------
// generic kernel: any (n_up, n_down) with K <= 4096, any strand.  One lane per centre.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS) scan_generic_kernel(
    const uint32_t *__restrict__ p2, const uint32_t *__restrict__ nmask, int64_t n_bases,
    const int64_t *__restrict__ chrom_off, const int64_t *__restrict__ chrom_len,
    const int32_t *__restrict__ reg_chrom, const int64_t *__restrict__ reg_start,
    const int64_t *__restrict__ reg_end, const int8_t *__restrict__ reg_strand, int64_t n_reg, int n_up,
    int n_down, int32_t *__restrict__ counts, unsigned long long *__restrict__ totals)
{
    extern __shared__ __align__(16) int smem[];
    const int klen = n_up + n_down + 1;
    const int K = 1 << (2 * klen);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    int *hist = smem + warp * K;
    for (int k = lane; k < K; k += 32) hist[k] = 0;
    __syncwarp();
    const int64_t gwarp = (int64_t)blockIdx.x * WARPS_PER_BLOCK + warp;
    const int64_t nwarps = (int64_t)gridDim.x * WARPS_PER_BLOCK;
    for (int64_t r = gwarp; r < n_reg; r += nwarps) {
        const bool minus = reg_strand != nullptr && __ldg(reg_strand + r) < 0;
        // on the minus strand the centre of the reverse-complemented k-mer has n_down bases to
        // its left and n_up to its right in plus-strand coordinates
        const int u = minus ? n_down : n_up;
        const int d = minus ? n_up : n_down;
        const RegionSpan sp = region_span(chrom_off, chrom_len, reg_chrom, reg_start, reg_end, r, n_up, n_down, u, d);
        for (int64_t g = sp.gs + lane; g < sp.ge; g += 32) {
            uint32_t key = 0;
            bool bad = false;
            for (int t = -u; t <= d; ++t) {
                const int64_t gg = g + t;
                if (gg < 0 || gg >= n_bases) {
                    bad = true;
                    continue;
                }
                const uint32_t code = (__ldg(p2 + (gg >> 4)) >> (30 - 2 * (int)(gg & 15))) & 3u;
                const uint32_t isn = (__ldg(nmask + (gg >> 5)) >> (31 - (int)(gg & 31))) & 1u;
                bad |= isn != 0u;
                key = (key << 2) | code;
            }
            if (!bad) atomicAdd(hist + (minus ? revcomp_key(key, klen) : key), 1);
        }
        __syncwarp();
        int32_t *out = counts + r * (int64_t)K;
        for (int k = lane; k < K; k += 32) {
            const int v = hist[k];
            hist[k] = 0;
            out[k] = v;
            if (totals != nullptr && v) atomicAdd(totals + k, (unsigned long long)v);
        }
        __syncwarp();
    }
}

template <int U>
int launch_sym(const uint32_t *p2, const uint32_t *nm, int64_t n_bases, const int64_t *chrom_off,
               const int64_t *chrom_len, const int32_t *reg_chrom, const int64_t *reg_start,
               const int64_t *reg_end, const int8_t *reg_strand, int64_t n_reg, int32_t *counts,
               unsigned long long *totals, cudaStream_t stream)
{
    constexpr int K = 1 << (2 * (2 * U + 1));
    const size_t smem = (size_t)WARPS_PER_BLOCK * K * sizeof(int) < 2048 ? 2048 : (size_t)WARPS_PER_BLOCK * K * sizeof(int);
    auto kern = scan_sym_kernel<U>;
    static thread_local int blocks_per_sm = 0;
    if (blocks_per_sm == 0) {
        DIG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DIG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, THREADS, smem));
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    int64_t blocks = (int64_t)dig::sm_count() * blocks_per_sm;
    const int64_t need = (n_reg + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    if (blocks > need) blocks = need;
    kern<<<(unsigned)blocks, THREADS, smem, stream>>>(reinterpret_cast<const uint2 *>(p2), p2, nm,
                                                      (n_bases + 31) >> 5, chrom_off, chrom_len, reg_chrom,
                                                      reg_start, reg_end, reg_strand, n_reg, counts, totals);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

}  // namespace

extern "C" int dig_count_contexts(const uint32_t *packed2_d, const uint32_t *nmask_d, int64_t n_bases,
                                  const int64_t *chrom_off_d, const int64_t *chrom_len_d,
                                  const int32_t *reg_chrom_d, const int64_t *reg_start_d,
                                  const int64_t *reg_end_d, const int8_t *reg_strand_d, int64_t n_reg, int n_up,
                                  int n_down, int32_t *counts_d, unsigned long long *totals_d, void *stream)
{
    DIG_CHECK_ARG(n_reg >= 0 && n_bases >= 0, "negative size");
    DIG_CHECK_ARG(n_up >= 0 && n_down >= 0 && n_up + n_down <= 5, "need 0 <= n_up, n_down and n_up + n_down <= 5");
    if (n_reg == 0) return DIG_OK;
    DIG_CHECK_ARG(packed2_d && nmask_d && chrom_off_d && chrom_len_d && reg_chrom_d && reg_start_d && reg_end_d &&
                      counts_d,
                  "null pointer");
    DIG_CHECK_ARG((reinterpret_cast<uintptr_t>(packed2_d) & 7u) == 0 && (reinterpret_cast<uintptr_t>(counts_d) & 15u) == 0,
                  "packed2_d must be 8-byte and counts_d 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_up == n_down && n_up <= 2) {
        switch (n_up) {
        case 0:
            return launch_sym<0>(packed2_d, nmask_d, n_bases, chrom_off_d, chrom_len_d, reg_chrom_d, reg_start_d,
                                 reg_end_d, reg_strand_d, n_reg, counts_d, totals_d, st);
        case 1:
            return launch_sym<1>(packed2_d, nmask_d, n_bases, chrom_off_d, chrom_len_d, reg_chrom_d, reg_start_d,
                                 reg_end_d, reg_strand_d, n_reg, counts_d, totals_d, st);
        default:
            return launch_sym<2>(packed2_d, nmask_d, n_bases, chrom_off_d, chrom_len_d, reg_chrom_d, reg_start_d,
                                 reg_end_d, reg_strand_d, n_reg, counts_d, totals_d, st);
        }
    }
    const int K = 1 << (2 * (n_up + n_down + 1));
    const size_t smem = (size_t)WARPS_PER_BLOCK * K * sizeof(int);
    static thread_local bool attr_set = false;
    if (!attr_set) {
        DIG_CUDA(cudaFuncSetAttribute(scan_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      WARPS_PER_BLOCK * 4096 * (int)sizeof(int)));
        attr_set = true;
    }
    int64_t blocks = (int64_t)dig::sm_count() * 4;
    const int64_t need = (n_reg + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    if (blocks > need) blocks = need;
    scan_generic_kernel<<<(unsigned)blocks, THREADS, smem, st>>>(packed2_d, nmask_d, n_bases, chrom_off_d,
                                                                 chrom_len_d, reg_chrom_d, reg_start_d, reg_end_d,
                                                                 reg_strand_d, n_reg, n_up, n_down, counts_d,
                                                                 totals_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}
