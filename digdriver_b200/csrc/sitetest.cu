// Per-site test (BASELINE.json config 3: ~30 M coding / splice sites, every site its own "element").
//
// In the reference a site set is an element whose L vector counts its sites per substitution
// (preprocess_sites, sequence_tools.py:647-711) and whose windows are those its sites fall in; nonc_model
// (genic_driver_tools.py:361-381) then gives P_SUM = sum_j (d_pr[j] / sum_i d_pr[i] R[i]) L[j].  For a set of ONE site,
// L is one-hot and R is the count row of the single window that contains the site:
//     P_site = d_pr[j] / denom[w, strand],   denom[w, +] = sum_i d_pr[i] R_w[i/3],   denom[w, -] uses R_w[revcomp(i/3)]
//     MU = Y_PRED[w], SIGMA = sqrt(STD[w]^2), ALPHA = MU^2/SIGMA^2, THETA = cj SIGMA^2/MU     (nb_model.py:237-241)
//     EXP = ALPHA THETA P,  PVAL = nb_pvalue_greater_midp(k, ALPHA, 1/(THETA P + 1))          (transfer_tools.py:343-355, :473-482)
// Running 30 M one-site elements through K6 would need a 192-double L row per site (46 GB); here the denominators
// are computed once per window (same lane assignment and reduction order as K6, so P is bit-identical to K6's) and
// every site is one thread.
#include "nb_math.cuh"

namespace {

using namespace dig_nb;

__device__ __forceinline__ int revcomp3(int c)
{
    const int b0 = (c >> 4) & 3, b1 = (c >> 2) & 3, b2 = c & 3;
    return ((3 - b2) << 4) | ((3 - b1) << 2) | (3 - b0);
}

// one warp per window; lane holds substitutions lane + 32 t like K6 (transfer.cu)
__global__ void __launch_bounds__(256) window_denominators_kernel(const int32_t *__restrict__ win_counts,
                                                                  const double *__restrict__ d_pr, int64_t n_win,
                                                                  double *__restrict__ denom_plus,
                                                                  double *__restrict__ denom_minus)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double dp[6];
#pragma unroll
    for (int t = 0; t < 6; ++t) dp[t] = __ldg(d_pr + lane + 32 * t);
    for (int64_t w = warp; w < n_win; w += nwarps) {
        const int32_t *row = win_counts + w * 64;
        double pp = 0.0, pm = 0.0;
#pragma unroll
        for (int t = 0; t < 6; ++t) {
            const int c = (lane + 32 * t) / 3;
            pp += dp[t] * (double)__ldg(row + c);
            pm += dp[t] * (double)__ldg(row + revcomp3(c));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            pp += __shfl_xor_sync(0xffffffffu, pp, o);
            pm += __shfl_xor_sync(0xffffffffu, pm, o);
        }
        if (lane == 0) {
            denom_plus[w] = pp;
            denom_minus[w] = pm;
        }
    }
}

__global__ void __launch_bounds__(256) site_test_kernel(
    const int32_t *__restrict__ site_chrom, const int64_t *__restrict__ site_start, const uint8_t *__restrict__ site_sub,
    const int8_t *__restrict__ site_strand, const double *__restrict__ site_k, int64_t n_site, int64_t window, int n_chrom,
    const int64_t *__restrict__ win_map_off, const int32_t *__restrict__ win_map, const double *__restrict__ y_pred,
    const double *__restrict__ stdv, const double *__restrict__ denom_plus, const double *__restrict__ denom_minus,
    const double *__restrict__ d_pr, double cj, double *__restrict__ p_out, double *__restrict__ exp_out,
    double *__restrict__ pval_out, int32_t *__restrict__ status)
{
    __shared__ double dp_s[192];
    for (int j = threadIdx.x; j < 192; j += blockDim.x) dp_s[j] = d_pr[j];
    __syncthreads();
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_site; i += stride) {
        const int32_t c = site_chrom[i];
        const int64_t wi = site_start[i] / window;                         // START >= 0
        const bool chrom_ok = c >= 0 && c < n_chrom;                       // outside the window map: KeyError below
        const int64_t m0 = chrom_ok ? __ldg(win_map_off + c) : 0, mn = chrom_ok ? __ldg(win_map_off + c + 1) - m0 : 0;
        const int32_t row = wi < mn ? __ldg(win_map + m0 + wi) : -1;
        if (row < 0) {                                                     // the reference raises KeyError here
            atomicMax(status, 2);
            if (p_out) p_out[i] = nan;
            if (exp_out) exp_out[i] = nan;
            pval_out[i] = nan;
            continue;
        }
        const double den = site_strand != nullptr && site_strand[i] < 0 ? __ldg(denom_minus + row) : __ldg(denom_plus + row);
        const double P = __ddiv_rn(dp_s[site_sub[i]], den);
        const double mu = __ldg(y_pred + row), sd = __ldg(stdv + row);
        const double var = __dmul_rn(sd, sd);
        const double sigma = sqrt(var);                                    // sqrt(sum of one sigma^2)
        const double s2 = __dmul_rn(sigma, sigma);
        const double alpha = __ddiv_rn(__dmul_rn(mu, mu), s2);
        const double theta = __dmul_rn(__ddiv_rn(s2, mu), cj);
        if (p_out) p_out[i] = P;
        if (exp_out) exp_out[i] = __dmul_rn(__dmul_rn(alpha, theta), P);
        const double p = __ddiv_rn(1.0, __dadd_rn(__dmul_rn(theta, P), 1.0));
        pval_out[i] = nb_midp(site_k[i], alpha, p);
    }
}

}  // namespace

extern "C" {

int dig_window_denominators(const int32_t *win_counts_d, const double *d_pr_d, int64_t n_win, double *denom_plus_d,
                            double *denom_minus_d, void *stream)
{
    DIG_CHECK_ARG(n_win >= 0, "negative size");
    if (n_win == 0) return DIG_OK;
    DIG_CHECK_ARG(win_counts_d && d_pr_d && denom_plus_d && denom_minus_d, "null pointer");
    int64_t blocks = (n_win * 32 + 255) / 256;
    const int64_t cap = (int64_t)dig::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    window_denominators_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(win_counts_d, d_pr_d, n_win,
                                                                                  denom_plus_d, denom_minus_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

int dig_site_test(const int32_t *site_chrom_d, const int64_t *site_start_d, const uint8_t *site_sub_d,
                  const int8_t *site_strand_d, const double *site_k_d, int64_t n_site, int64_t window, int n_chrom,
                  const int64_t *win_map_off_d, const int32_t *win_map_d, const double *y_pred_d, const double *std_d,
                  const double *denom_plus_d, const double *denom_minus_d, const double *d_pr_d, double cj,
                  double *p_out_d, double *exp_out_d, double *pval_out_d, int32_t *status_d, void *stream)
{
    DIG_CHECK_ARG(n_site >= 0 && window > 0 && n_chrom >= 0, "bad sizes");
    DIG_CHECK_ARG(status_d != nullptr, "null pointer");
    DIG_CUDA(cudaMemsetAsync(status_d, 0, sizeof(int32_t), (cudaStream_t)stream));
    if (n_site == 0) return DIG_OK;
    DIG_CHECK_ARG(site_chrom_d && site_start_d && site_sub_d && site_k_d && win_map_off_d && win_map_d && y_pred_d &&
                      std_d && denom_plus_d && denom_minus_d && d_pr_d && pval_out_d,
                  "null pointer");
    int64_t blocks = (n_site + 255) / 256;
    const int64_t cap = (int64_t)dig::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    site_test_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        site_chrom_d, site_start_d, site_sub_d, site_strand_d, site_k_d, n_site, window, n_chrom, win_map_off_d, win_map_d,
        y_pred_d, std_d, denom_plus_d, denom_minus_d, d_pr_d, cj, p_out_d, exp_out_d, pval_out_d, status_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

}
