// K6: element transfer -- overlapped-window enumeration, region parameters (mu, sigma, R_obs,
// flag) and the context-weighted mutability fraction P = sum_j (d_pr[j] / sum_i d_pr[i] R[i]) L[j].
//
// One warp per element.  The set of windows an element overlaps (the Python `set` of
// get_ideal_overlaps, genic_driver_tools.py:275-283) is a bitmap over the element's window span
// in shared memory: blocks OR their ranges in, then the warp walks the set bits in ascending
// order, gathering each window's 64 trinucleotide counts with one coalesced 256-byte load and
// the per-cohort region parameters.  Sums therefore have a fixed, deterministic order
// (ascending window), unlike the hash order of the reference's set.
#include "dig_common.cuh"

namespace {

constexpr int TW = 4;                  // warps per block
constexpr int MAX_COHORT_PER_LANE = 2; // n_cohort <= 64 per call

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// trinucleotide index -> index of its reverse complement
__device__ __forceinline__ int revcomp3(int c)
{
    const int b0 = (c >> 4) & 3, b1 = (c >> 2) & 3, b2 = c & 3;
    return ((3 - b2) << 4) | ((3 - b1) << 2) | (3 - b0);
}

__device__ __forceinline__ int64_t floor_div(int64_t a, int64_t b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
__device__ __forceinline__ int64_t ceil_div(int64_t a, int64_t b) { return a >= 0 ? (a + b - 1) / b : -((-a) / b); }

// RC = true (dig_element_region_counts) also writes the strand-flipped 64-context region counts of every element and
// tolerates n_cohort == 0 with the cohort / result pointers NULL; RC = false is the production K6 instantiation.
template <bool RC>
__global__ void __launch_bounds__(TW * 32, 8) transfer_kernel(
    const int32_t *__restrict__ elt_chrom, const int8_t *__restrict__ elt_strand,
    const int64_t *__restrict__ blk_ptr, const int64_t *__restrict__ blk_start,
    const int64_t *__restrict__ blk_end, int64_t n_elt, int64_t window, int n_chrom,
    const int64_t *__restrict__ win_map_off, const int32_t *__restrict__ win_map, const int32_t *__restrict__ win_counts,
    const double *__restrict__ y_pred, const double *__restrict__ stdv, const double *__restrict__ y_true,
    const uint8_t *__restrict__ flag, int64_t n_win, int n_cohort, const double *__restrict__ d_pr,
    const int32_t *__restrict__ blk_counts, const double *__restrict__ L_elt, int n_col, int span_words,
    double *__restrict__ mu_out, double *__restrict__ sigma_out, double *__restrict__ robs_out,
    uint8_t *__restrict__ flag_out, int64_t *__restrict__ r_size, int64_t *__restrict__ elt_size,
    double *__restrict__ p_out, int32_t *__restrict__ n_win_out, int32_t *__restrict__ status,
    int64_t *__restrict__ region_counts_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // per-warp layout: r64[64] double | l64[64] double | bitmap[span_words] uint32
    const size_t per_warp = 128 * sizeof(double) + (size_t)span_words * sizeof(uint32_t);
    double *r64 = reinterpret_cast<double *>(smem_raw + warp * per_warp);
    double *l64 = r64 + 64;
    uint32_t *bitmap = reinterpret_cast<uint32_t *>(l64 + 64);

    const bool l4_aligned = (reinterpret_cast<uintptr_t>(L_elt) & 15u) == 0;
    const int64_t gwarp = (int64_t)blockIdx.x * TW + warp;
    const int64_t nwarps = (int64_t)gridDim.x * TW;

    for (int64_t e = gwarp; e < n_elt; e += nwarps) {
        const int64_t b0 = blk_ptr[e], b1 = blk_ptr[e + 1];
        const int32_t c = elt_chrom[e];
        const bool minus = elt_strand[e] < 0;
        // ---- window span of the element
        int64_t wmin = INT64_MAX, wmax = INT64_MIN;
        for (int64_t b = b0 + lane; b < b1; b += 32) {
            const int64_t lo = floor_div(blk_start[b], window);
            const int64_t hi = ceil_div(blk_end[b], window);          // windows [lo, hi)
            if (hi > lo) {
                wmin = lo < wmin ? lo : wmin;
                wmax = hi > wmax ? hi : wmax;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int64_t a = __shfl_xor_sync(0xffffffffu, wmin, o);
            const int64_t b = __shfl_xor_sync(0xffffffffu, wmax, o);
            wmin = a < wmin ? a : wmin;
            wmax = b > wmax ? b : wmax;
        }
        const int64_t span = wmax > wmin ? wmax - wmin : 0;
        const int words = (int)((span + 31) >> 5);
        bool ok = true;
        if (words > span_words) {
            if (lane == 0) atomicMax(status, 3);
            ok = false;
        }
        // ---- bitmap of overlapped windows
        if (ok) {
            for (int w = lane; w < words; w += 32) bitmap[w] = 0u;
            __syncwarp();
            for (int64_t b = b0 + lane; b < b1; b += 32) {
                const int64_t lo = floor_div(blk_start[b], window) - wmin;
                const int64_t hi = ceil_div(blk_end[b], window) - wmin;
                for (int64_t w = lo; w < hi;) {
                    const int word = (int)(w >> 5);
                    const int bit = (int)(w & 31);
                    const int64_t upto = ((int64_t)(word + 1) << 5) < hi ? ((int64_t)(word + 1) << 5) : hi;
                    const int nbits = (int)(upto - w);
                    const uint32_t m = (nbits >= 32 ? 0xFFFFFFFFu : ((1u << nbits) - 1u)) << bit;
                    atomicOr(bitmap + word, m);
                    w = upto;
                }
            }
            __syncwarp();
        }
        // ---- gather window counts and region parameters in ascending window order
        double r_lo = 0.0, r_hi = 0.0;                       // bins lane and lane+32 (exact integers)
        double mu[MAX_COHORT_PER_LANE], var[MAX_COHORT_PER_LANE], ro[MAX_COHORT_PER_LANE];
        int fl[MAX_COHORT_PER_LANE];
#pragma unroll
        for (int q = 0; q < MAX_COHORT_PER_LANE; ++q) mu[q] = var[q] = ro[q] = 0.0, fl[q] = 0;
        int nw = 0;
        // a chromosome outside the window map has no window at all: every lookup below misses and the element gets
        // status 2 (the reference raises KeyError for it)
        const bool chrom_ok = c >= 0 && c < n_chrom;
        const int64_t map0 = chrom_ok ? win_map_off[c] : 0;
        const int64_t map_n = chrom_ok ? win_map_off[c + 1] - map0 : 0;
        for (int w = 0; ok && w < words; ++w) {
            const uint32_t bits = bitmap[w];
            const int cnt = __popc(bits);
            if (cnt == 0) continue;
            // the rows of all (up to 32) set windows of this word are looked up by 32 lanes at once; the gather loop
            // below then has no load that depends on a previous load, so several windows are in flight together
            int32_t my_row = -1;
            if (lane < cnt) {
                const int64_t wi = wmin + ((int64_t)w << 5) + (int64_t)__fns(bits, 0, lane + 1);
                my_row = (wi >= 0 && wi < map_n) ? __ldg(win_map + map0 + wi) : -1;
            }
            if (__any_sync(0xffffffffu, lane < cnt && my_row < 0) && lane == 0) atomicMax(status, 2);   // KeyError in the reference
            // counts: four windows' rows in flight, added in ascending window order
            for (int k0 = 0; k0 < cnt; k0 += 4) {
                int32_t row[4], clo[4], chi[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    row[u] = k0 + u < cnt ? __shfl_sync(0xffffffffu, my_row, (k0 + u) & 31) : -1;
                    clo[u] = chi[u] = 0;
                    if (row[u] >= 0) {
                        const int32_t *wc = win_counts + (int64_t)row[u] * 64;
                        clo[u] = __ldg(wc + lane);
                        chi[u] = __ldg(wc + lane + 32);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (row[u] < 0) continue;
                    ++nw;
                    r_lo += (double)clo[u];
                    r_hi += (double)chi[u];
                }
            }
            if (n_cohort == 1) {
                // one cohort (the CLI / bench case): lane k loads the region parameters of window k, so all windows of
                // this word are fetched at once; the sums are then formed in ascending window order through shuffles
                double yp = 0.0, sd = 0.0, yt = 0.0;
                int fg = 0;
                if (lane < cnt && my_row >= 0) {
                    yp = __ldg(y_pred + my_row);
                    sd = __ldg(stdv + my_row);
                    yt = __ldg(y_true + my_row);
                    fg = __ldg(flag + my_row) != 0;
                }
                for (int k = 0; k < cnt; ++k) {
                    const double a = __shfl_sync(0xffffffffu, yp, k), b = __shfl_sync(0xffffffffu, sd, k);
                    const double t = __shfl_sync(0xffffffffu, yt, k);
                    const int f = __shfl_sync(0xffffffffu, fg, k);
                    if (__shfl_sync(0xffffffffu, my_row, k) < 0) continue;
                    mu[0] += a;
                    var[0] = __dadd_rn(var[0], __dmul_rn(b, b));
                    ro[0] += t;
                    fl[0] |= f;
                }
            } else {
                for (int k = 0; k < cnt; ++k) {
                    const int32_t row = __shfl_sync(0xffffffffu, my_row, k);
                    if (row < 0) continue;
#pragma unroll
                    for (int q = 0; q < MAX_COHORT_PER_LANE; ++q) {
                        const int ci = lane + 32 * q;
                        if (ci < n_cohort) {
                            const int64_t o = (int64_t)ci * n_win + row;
                            const double s = __ldg(stdv + o);
                            mu[q] += __ldg(y_pred + o);
                            var[q] = __dadd_rn(var[q], __dmul_rn(s, s));
                            ro[q] += __ldg(y_true + o);
                            fl[q] |= __ldg(flag + o) != 0;
                        }
                    }
                }
            }
        }
        // ---- strand: new[ctx] = old[revcomp(ctx)] (sequence_tools.py:633-634)
        __syncwarp();
        if (!minus) {
            r64[lane] = r_lo;
            r64[lane + 32] = r_hi;
        } else {
            r64[revcomp3(lane)] = r_lo;
            r64[revcomp3(lane + 32)] = r_hi;
        }
        // ---- L in 64-context form (element mode): sum of the element's block counts
        double l_lo = 0.0, l_hi = 0.0;
        int64_t gene_len = 0;
        if (blk_counts != nullptr) {
            for (int64_t b = b0; b < b1; ++b) {
                l_lo += (double)__ldg(blk_counts + b * 64 + lane);
                l_hi += (double)__ldg(blk_counts + b * 64 + lane + 32);
            }
            l64[lane] = l_lo;
            l64[lane + 32] = l_hi;
        } else {
            for (int64_t b = b0 + lane; b < b1; b += 32) gene_len += blk_end[b] - blk_start[b] + 1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) gene_len += __shfl_xor_sync(0xffffffffu, gene_len, o);
        }
        __syncwarp();
        if (RC) {
            region_counts_out[e * 64 + lane] = (int64_t)r64[lane];
            region_counts_out[e * 64 + lane + 32] = (int64_t)r64[lane + 32];
        }
        const double rsum = warp_sum(r_lo + r_hi);
        const double lsum = warp_sum(l_lo + l_hi);

        // ---- per cohort: denom = sum_j d_pr[j] R192[j];  P[col] = sum_j (d_pr[j]/denom) L[j][col]
        for (int ci = 0; ci < n_cohort; ++ci) {
            const double *dp = d_pr + (int64_t)ci * 192;
            double part = 0.0;
#pragma unroll
            for (int t = 0; t < 6; ++t) {
                const int j = lane + 32 * t;
                part += __ldg(dp + j) * r64[j / 3];
            }
            const double denom = warp_sum(part);
            double wgt[6];                                    // d_pr[j] / denom, shared by every column
#pragma unroll
            for (int t = 0; t < 6; ++t) wgt[t] = __ldg(dp + lane + 32 * t) / denom;
            if (blk_counts == nullptr && n_col == 4 && l4_aligned) {
                // gene mode: the four columns of a substitution are 32 contiguous bytes -- two 128-bit loads per
                // substitution, all twelve issued before the first use (same products and summation order per column)
                const double2 *L2 = reinterpret_cast<const double2 *>(L_elt + (int64_t)e * 192 * 4);
                double2 la[6], lb[6];
#pragma unroll
                for (int t = 0; t < 6; ++t) {
                    la[t] = __ldg(L2 + 2 * (lane + 32 * t));
                    lb[t] = __ldg(L2 + 2 * (lane + 32 * t) + 1);
                }
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                for (int t = 0; t < 6; ++t) {
                    a0 += wgt[t] * la[t].x;
                    a1 += wgt[t] * la[t].y;
                    a2 += wgt[t] * lb[t].x;
                    a3 += wgt[t] * lb[t].y;
                }
                a0 = warp_sum(a0);
                a1 = warp_sum(a1);
                a2 = warp_sum(a2);
                a3 = warp_sum(a3);
                if (lane == 0) {
                    double *po = p_out + ((int64_t)ci * n_elt + e) * 4;
                    po[0] = a0;
                    po[1] = a1;
                    po[2] = a2;
                    po[3] = a3;
                }
            } else {
                for (int col = 0; col < n_col; ++col) {
                    double acc = 0.0;
#pragma unroll
                    for (int t = 0; t < 6; ++t) {
                        const int j = lane + 32 * t;
                        const double Lj = blk_counts != nullptr ? l64[j / 3]
                                                                : __ldg(L_elt + ((int64_t)e * 192 + j) * n_col + col);
                        acc += wgt[t] * Lj;
                    }
                    acc = warp_sum(acc);
                    if (lane == 0) p_out[((int64_t)ci * n_elt + e) * n_col + col] = acc;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < MAX_COHORT_PER_LANE; ++q) {
            const int ci = lane + 32 * q;
            if (ci < n_cohort) {
                const int64_t o = (int64_t)ci * n_elt + e;
                mu_out[o] = mu[q];
                sigma_out[o] = sqrt(var[q]);
                robs_out[o] = ro[q];
                flag_out[o] = (uint8_t)fl[q];
            }
        }
        if (lane == 0) {
            if (!RC || r_size) r_size[e] = (int64_t)rsum;    // int(sum(R192)/3) == sum(R64)
            if (!RC || elt_size) elt_size[e] = blk_counts != nullptr ? (int64_t)lsum : gene_len;
            n_win_out[e] = nw;
        }
        __syncwarp();
    }
}

}  // namespace

extern "C" int dig_element_transfer(const int32_t *elt_chrom_d, const int8_t *elt_strand_d, const int64_t *blk_ptr_d,
                                    const int64_t *blk_start_d, const int64_t *blk_end_d, int64_t n_elt,
                                    int64_t window, int n_chrom, const int64_t *win_map_off_d, const int32_t *win_map_d,
                                    const int32_t *win_counts_d, const double *y_pred_d, const double *std_d,
                                    const double *y_true_d, const uint8_t *flag_d, int64_t n_win, int n_cohort,
                                    const double *d_pr_d, const int32_t *blk_counts_d, const double *L_elt_d,
                                    int n_col, int max_span_windows, double *mu_d, double *sigma_d, double *r_obs_d,
                                    uint8_t *flag_out_d, int64_t *r_size_d, int64_t *elt_size_d, double *p_out_d,
                                    int32_t *n_win_out_d, int32_t *status_d, void *stream)
{
    DIG_CHECK_ARG(n_elt >= 0 && n_win >= 0 && window > 0 && n_chrom >= 0, "bad sizes");
    DIG_CHECK_ARG(n_cohort >= 1 && n_cohort <= 32 * MAX_COHORT_PER_LANE, "n_cohort must be in [1, 64] per call");
    DIG_CHECK_ARG((blk_counts_d != nullptr) != (L_elt_d != nullptr), "pass exactly one of blk_counts_d / L_elt_d");
    DIG_CHECK_ARG(blk_counts_d == nullptr || n_col == 1, "n_col must be 1 with blk_counts_d");
    DIG_CHECK_ARG(n_col >= 1 && max_span_windows >= 1, "bad n_col / max_span_windows");
    DIG_CHECK_ARG(status_d != nullptr, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    DIG_CUDA(cudaMemsetAsync(status_d, 0, sizeof(int32_t), st));
    if (n_elt == 0) return DIG_OK;
    DIG_CHECK_ARG(elt_chrom_d && elt_strand_d && blk_ptr_d && blk_start_d && blk_end_d && win_map_off_d &&
                      win_map_d && win_counts_d && y_pred_d && std_d && y_true_d && flag_d && d_pr_d && mu_d &&
                      sigma_d && r_obs_d && flag_out_d && r_size_d && elt_size_d && p_out_d && n_win_out_d,
                  "null pointer");
    const int span_words = ((max_span_windows + 31) / 32 + 3) & ~3;      // keeps every warp's slice 16-byte aligned
    const size_t smem = (size_t)TW * (128 * sizeof(double) + (size_t)span_words * sizeof(uint32_t));
    if (smem > 200 * 1024) {
        dig::set_error("dig_element_transfer: window span of %d windows needs %zu B of shared memory",
                       max_span_windows, smem);
        return DIG_ERR_UNSUPPORTED;
    }
    DIG_CUDA(cudaFuncSetAttribute(transfer_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (n_elt + TW - 1) / TW;
    const int64_t cap = (int64_t)dig::sm_count() * 64;      // ~one element per warp: the hardware scheduler balances long genes
    if (blocks > cap) blocks = cap;
    transfer_kernel<false><<<(unsigned)blocks, TW * 32, smem, st>>>(
        elt_chrom_d, elt_strand_d, blk_ptr_d, blk_start_d, blk_end_d, n_elt, window, n_chrom, win_map_off_d, win_map_d,
        win_counts_d, y_pred_d, std_d, y_true_d, flag_d, n_win, n_cohort, d_pr_d, blk_counts_d, L_elt_d, n_col,
        span_words, mu_d, sigma_d, r_obs_d, flag_out_d, r_size_d, elt_size_d, p_out_d, n_win_out_d, status_d, nullptr);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

// The persisted intermediate of preprocess_nonc / preprocess_sites (sequence_tools.py:596-711): region_counts of every
// element = sum of the 64 trinucleotide counts of its overlapped windows, reverse-complemented for minus-strand
// elements (the reference stores it repeated x3 as 192 values).  Same kernel as dig_element_transfer, no cohorts.
extern "C" int dig_element_region_counts(const int32_t *elt_chrom_d, const int8_t *elt_strand_d, const int64_t *blk_ptr_d,
                                         const int64_t *blk_start_d, const int64_t *blk_end_d, int64_t n_elt,
                                         int64_t window, int n_chrom, const int64_t *win_map_off_d, const int32_t *win_map_d,
                                         const int32_t *win_counts_d, int64_t n_win, int max_span_windows,
                                         int64_t *region_counts_d, int32_t *n_win_out_d, int32_t *status_d, void *stream)
{
    DIG_CHECK_ARG(n_elt >= 0 && n_win >= 0 && window > 0 && max_span_windows >= 1 && n_chrom >= 0, "bad sizes");
    DIG_CHECK_ARG(status_d != nullptr, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    DIG_CUDA(cudaMemsetAsync(status_d, 0, sizeof(int32_t), st));
    if (n_elt == 0) return DIG_OK;
    DIG_CHECK_ARG(elt_chrom_d && elt_strand_d && blk_ptr_d && blk_start_d && blk_end_d && win_map_off_d && win_map_d &&
                      win_counts_d && region_counts_d && n_win_out_d,
                  "null pointer");
    const int span_words = ((max_span_windows + 31) / 32 + 3) & ~3;
    const size_t smem = (size_t)TW * (128 * sizeof(double) + (size_t)span_words * sizeof(uint32_t));
    if (smem > 200 * 1024) {
        dig::set_error("dig_element_region_counts: window span of %d windows needs %zu B of shared memory",
                       max_span_windows, smem);
        return DIG_ERR_UNSUPPORTED;
    }
    DIG_CUDA(cudaFuncSetAttribute(transfer_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (n_elt + TW - 1) / TW;
    const int64_t cap = (int64_t)dig::sm_count() * 64;
    if (blocks > cap) blocks = cap;
    transfer_kernel<true><<<(unsigned)blocks, TW * 32, smem, st>>>(
        elt_chrom_d, elt_strand_d, blk_ptr_d, blk_start_d, blk_end_d, n_elt, window, n_chrom, win_map_off_d, win_map_d,
        win_counts_d, nullptr, nullptr, nullptr, nullptr, n_win, 0, nullptr, nullptr, nullptr, 0, span_words, nullptr,
        nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, n_win_out_d, status_d, region_counts_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

// P_SUM of nonc_model (genic_driver_tools.py:361-369) from the PERSISTED intermediates: prob_sum = region_counts * d_pr,
// t_pi = d_pr / prob_sum.sum(), p_mut = (t_pi * L).sum().  One warp per element, the same lane assignment and
// summation order as transfer_kernel, so the result is bit-identical to dig_element_transfer's P for the same counts.
namespace {

__global__ void __launch_bounds__(128) element_psum_kernel(const double *__restrict__ L, const int64_t *__restrict__ R,
                                                           const double *__restrict__ d_pr, int64_t n_elt,
                                                           double *__restrict__ p_out, double *__restrict__ denom_out)
{
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = gwarp; e < n_elt; e += nwarps) {
        double part = 0.0;
#pragma unroll
        for (int t = 0; t < 6; ++t) {
            const int j = lane + 32 * t;
            part += __ldg(d_pr + j) * (double)__ldg(R + e * 192 + j);
        }
        const double denom = warp_sum(part);
        double acc = 0.0;
#pragma unroll
        for (int t = 0; t < 6; ++t) {
            const int j = lane + 32 * t;
            acc += (__ldg(d_pr + j) / denom) * __ldg(L + e * 192 + j);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            p_out[e] = acc;
            if (denom_out) denom_out[e] = denom;
        }
    }
}

}  // namespace

extern "C" int dig_element_psum(const double *L_d, const int64_t *region_counts_d, const double *d_pr_d, int64_t n_elt,
                                double *p_out_d, double *denom_out_d, void *stream)
{
    DIG_CHECK_ARG(n_elt >= 0, "negative size");
    if (n_elt == 0) return DIG_OK;
    DIG_CHECK_ARG(L_d && region_counts_d && d_pr_d && p_out_d, "null pointer");
    int64_t blocks = (n_elt + 3) / 4;
    const int64_t cap = (int64_t)dig::sm_count() * 32;
    if (blocks > cap) blocks = cap;
    element_psum_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(L_d, region_counts_d, d_pr_d, n_elt, p_out_d,
                                                                           denom_out_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}
