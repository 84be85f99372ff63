// K8: per-position (or per-bin) negative-binomial hotspot test over a list of regions.
//
// Replaces the per-base Python loops of
//   base_probabilities_by_region (sequence_tools.py:292-317): pt(pos) = S_prob[k-mer at pos] / sum over the region,
//       0 for a k-mer that contains N;
//   apply_nb_to_region (nb_model.py:126-186): k(pos) = number of mutation rows whose START is pos, p = 1/(pt theta + 1),
//       pval = nb_pvalue_exact(k, alpha, p) (nb_model.py:298-314), exp = pt mu; with binsize > 1 consecutive positions
//       are pooled (pt and k summed, POS = mean position);
//   nb_model (nb_model.py:188-235) which concatenates the regions.
// The positions of a region are exactly the centres the context scan walks (region_span in scan_common.cuh), so the
// normaliser is the dot product of the region's K2 count row with S_prob.
//
//   dig_region_prob_norm : norm[r] = sum_k counts[r, k] * S_prob[k]                 (one warp per region, FP64)
//   dig_position_obs     : obs[bin] = mutations with START inside the bin            (one warp per region, sorted keys)
//   dig_position_test    : pt, exp, pval (and POS) per bin                           (one CTA per region, S_prob in smem)
#include "nb_math.cuh"
#include "scan_common.cuh"

namespace {

using namespace dig_nb;
using namespace digscan;

__global__ void __launch_bounds__(256) region_prob_norm_kernel(const int32_t *__restrict__ counts,
                                                               const double *__restrict__ s_prob, int64_t n_reg, int K,
                                                               double *__restrict__ norm)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_reg; r += nwarps) {
        const int32_t *row = counts + r * (int64_t)K;
        double acc = 0.0;
        for (int k = lane; k < K; k += 32) acc += (double)row[k] * s_prob[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) norm[r] = acc;
    }
}

__device__ __forceinline__ int64_t lower_bound_key(const int64_t *__restrict__ a, int64_t n, int64_t key)
{
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// mut_key = chrom index << 32 | START, ascending.  One warp per region.
__global__ void __launch_bounds__(256) position_obs_kernel(
    const int64_t *__restrict__ mut_key, int64_t n_mut, const int64_t *__restrict__ chrom_off,
    const int64_t *__restrict__ chrom_len, const int32_t *__restrict__ reg_chrom, const int64_t *__restrict__ reg_start,
    const int64_t *__restrict__ reg_end, int64_t n_reg, int n_up, int n_down, int binsize,
    const int64_t *__restrict__ bin_ptr, int32_t *__restrict__ obs)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_reg; r += nwarps) {
        const RegionSpan sp = region_span(chrom_off, chrom_len, reg_chrom, reg_start, reg_end, r, n_up, n_down, n_up, n_down);
        if (sp.ge <= sp.gs) continue;
        const int32_t c = __ldg(reg_chrom + r);
        const int64_t off = __ldg(chrom_off + c);
        const int64_t p0 = sp.gs - off, p1 = sp.ge - off;                    // chromosome coordinates of the centres
        const int64_t lo = lower_bound_key(mut_key, n_mut, ((int64_t)c << 32) | p0);
        const int64_t hi = lower_bound_key(mut_key, n_mut, ((int64_t)c << 32) | p1);
        const int64_t b0 = __ldg(bin_ptr + r);
        for (int64_t i = lo + lane; i < hi; i += 32) {
            const int64_t pos = __ldg(mut_key + i) & 0xFFFFFFFFll;
            atomicAdd(obs + b0 + (pos - p0) / binsize, 1);
        }
    }
}

// S_prob of the k-mer centred on global position g, 0 when it touches a non-ACGT base.  The k-mer (at most 6 bases)
// lies inside two consecutive packed words and its N bits inside two consecutive mask words: four loads, one 64-bit
// shift each.  last_p / last_m are the last valid word indices: a k-mer that ends in the last word never needs the
// (clamped) second word.
__device__ __forceinline__ double kmer_prob(const uint32_t *__restrict__ p2, const uint32_t *__restrict__ nmask,
                                            int64_t g, int n_up, int klen, int64_t last_p, int64_t last_m,
                                            const double *s_prob_s)
{
    const int64_t f = g - n_up;                                   // first base of the k-mer
    const int64_t w = f >> 4;
    const unsigned long long pk = ((unsigned long long)__ldg(p2 + w) << 32) | __ldg(p2 + (w < last_p ? w + 1 : last_p));
    const uint32_t key = (uint32_t)(pk >> (64 - 2 * ((int)(f & 15) + klen))) & ((1u << (2 * klen)) - 1u);
    const int64_t m = f >> 5;
    const unsigned long long mk = ((unsigned long long)__ldg(nmask + m) << 32) | __ldg(nmask + (m < last_m ? m + 1 : last_m));
    const uint32_t bad = (uint32_t)(mk >> (64 - ((int)(f & 31) + klen))) & ((1u << klen) - 1u);
    return bad ? 0.0 : s_prob_s[key];
}

__global__ void __launch_bounds__(256, 4) position_test_kernel(
    const uint32_t *__restrict__ p2, const uint32_t *__restrict__ nmask, int64_t last_p, int64_t last_m,
    const int64_t *__restrict__ chrom_off,
    const int64_t *__restrict__ chrom_len, const int32_t *__restrict__ reg_chrom, const int64_t *__restrict__ reg_start,
    const int64_t *__restrict__ reg_end, int64_t n_reg, int n_up, int n_down, const double *__restrict__ s_prob,
    const double *__restrict__ norm, const double *__restrict__ mu, const double *__restrict__ sigma, int binsize,
    const int64_t *__restrict__ bin_ptr, const int32_t *__restrict__ obs, double *__restrict__ pval,
    double *__restrict__ pt_out, double *__restrict__ exp_out, double *__restrict__ pos_out)
{
    extern __shared__ double s_prob_s[];
    const int klen = n_up + n_down + 1;
    const int K = 1 << (2 * klen);
    for (int k = threadIdx.x; k < K; k += blockDim.x) s_prob_s[k] = s_prob[k];
    __syncthreads();
    for (int64_t r = blockIdx.x; r < n_reg; r += gridDim.x) {
        const RegionSpan sp = region_span(chrom_off, chrom_len, reg_chrom, reg_start, reg_end, r, n_up, n_down, n_up, n_down);
        const int64_t n_pos = sp.ge - sp.gs;
        if (n_pos <= 0) continue;
        const int64_t off = __ldg(chrom_off + __ldg(reg_chrom + r));
        const int64_t n_bin = (n_pos + binsize - 1) / binsize;
        const int64_t b0 = __ldg(bin_ptr + r);
        const double m = mu[r], s = sigma[r];
        const double alpha = (m * m) / (s * s);                 // normal_params_to_gamma (nb_model.py:237-241)
        const double theta = (s * s) / m;
        const double nrm = norm != nullptr ? norm[r] : 1.0;
        for (int64_t b = threadIdx.x; b < n_bin; b += blockDim.x) {
            const int64_t g0 = sp.gs + b * binsize;
            const int64_t g1 = g0 + binsize < sp.ge ? g0 + binsize : sp.ge;
            double pt = 0.0;
            // the reference normalises every position first and then sums the bin (nb_model.py:165)
            for (int64_t g = g0; g < g1; ++g) pt += __ddiv_rn(kmer_prob(p2, nmask, g, n_up, klen, last_p, last_m, s_prob_s), nrm);
            const double k = (double)obs[b0 + b];
            const double p = __ddiv_rn(1.0, __dadd_rn(__dmul_rn(pt, theta), 1.0));
            pval[b0 + b] = nb_exact(k, alpha, p);
            if (pt_out) pt_out[b0 + b] = pt;
            if (exp_out) exp_out[b0 + b] = __dmul_rn(pt, m);
            // np.mean of the integer positions of the bin (chromosome coordinates)
            if (pos_out) pos_out[b0 + b] = (double)(g0 - off) + 0.5 * (double)(g1 - g0 - 1);
        }
    }
}

inline unsigned grid_warps(int64_t n_reg, int threads)
{
    int64_t blocks = (n_reg * 32 + threads - 1) / threads;
    const int64_t cap = (int64_t)dig::sm_count() * 8;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}

}  // namespace

extern "C" {

int dig_region_prob_norm(const int32_t *counts_d, const double *s_prob_d, int64_t n_reg, int n_ctx, double *norm_d,
                         void *stream)
{
    DIG_CHECK_ARG(n_reg >= 0 && n_ctx > 0, "bad size");
    if (n_reg == 0) return DIG_OK;
    DIG_CHECK_ARG(counts_d && s_prob_d && norm_d, "null pointer");
    region_prob_norm_kernel<<<grid_warps(n_reg, 256), 256, 0, (cudaStream_t)stream>>>(counts_d, s_prob_d, n_reg, n_ctx,
                                                                                        norm_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

int dig_position_obs(const int64_t *mut_key_d, int64_t n_mut, const int64_t *chrom_off_d, const int64_t *chrom_len_d,
                     const int32_t *reg_chrom_d, const int64_t *reg_start_d, const int64_t *reg_end_d, int64_t n_reg,
                     int n_up, int n_down, int binsize, const int64_t *bin_ptr_d, int64_t n_bin, int32_t *obs_d,
                     void *stream)
{
    DIG_CHECK_ARG(n_reg >= 0 && n_mut >= 0 && n_bin >= 0 && binsize >= 1, "bad size");
    DIG_CHECK_ARG(n_up >= 0 && n_down >= 0 && n_up + n_down <= 5, "need 0 <= n_up, n_down and n_up + n_down <= 5");
    if (n_bin > 0) {
        DIG_CHECK_ARG(obs_d, "null pointer");
        DIG_CUDA(cudaMemsetAsync(obs_d, 0, (size_t)n_bin * sizeof(int32_t), (cudaStream_t)stream));
    }
    if (n_reg == 0 || n_mut == 0) return DIG_OK;
    DIG_CHECK_ARG(mut_key_d && chrom_off_d && chrom_len_d && reg_chrom_d && reg_start_d && reg_end_d && bin_ptr_d,
                  "null pointer");
    position_obs_kernel<<<grid_warps(n_reg, 256), 256, 0, (cudaStream_t)stream>>>(
        mut_key_d, n_mut, chrom_off_d, chrom_len_d, reg_chrom_d, reg_start_d, reg_end_d, n_reg, n_up, n_down, binsize,
        bin_ptr_d, obs_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

int dig_position_test(const uint32_t *packed2_d, const uint32_t *nmask_d, int64_t n_bases, const int64_t *chrom_off_d,
                      const int64_t *chrom_len_d, const int32_t *reg_chrom_d, const int64_t *reg_start_d,
                      const int64_t *reg_end_d, int64_t n_reg, int n_up, int n_down, const double *s_prob_d,
                      const double *norm_d, const double *mu_d, const double *sigma_d, int binsize,
                      const int64_t *bin_ptr_d, const int32_t *obs_d, double *pval_d, double *pt_d, double *exp_d,
                      double *pos_d, void *stream)
{
    DIG_CHECK_ARG(n_reg >= 0 && n_bases >= 0 && binsize >= 1, "bad size");
    DIG_CHECK_ARG(n_up >= 0 && n_down >= 0 && n_up + n_down <= 5, "need 0 <= n_up, n_down and n_up + n_down <= 5");
    if (n_reg == 0) return DIG_OK;
    DIG_CHECK_ARG(packed2_d && nmask_d && chrom_off_d && chrom_len_d && reg_chrom_d && reg_start_d && reg_end_d &&
                      s_prob_d && mu_d && sigma_d && bin_ptr_d && obs_d && pval_d,
                  "null pointer");
    const int K = 1 << (2 * (n_up + n_down + 1));
    const size_t smem = (size_t)K * sizeof(double);
    DIG_CUDA(cudaFuncSetAttribute(position_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 8));
    int64_t blocks = (int64_t)dig::sm_count() * 8;
    if (blocks > n_reg) blocks = n_reg;
    position_test_kernel<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(
        packed2_d, nmask_d, (((n_bases + 31) >> 5) << 1) - 1, ((n_bases + 31) >> 5) - 1, chrom_off_d, chrom_len_d,
        reg_chrom_d, reg_start_d, reg_end_d, n_reg, n_up, n_down, s_prob_d, norm_d, mu_d, sigma_d, binsize, bin_ptr_d, obs_d, pval_d, pt_d, exp_d, pos_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

int dig_nb_pvalue_exact(const double *k_d, const double *alpha_d, const double *p_d, int64_t n, double *pval_out_d,
                        void *stream);

}

namespace {
__global__ void __launch_bounds__(128) nb_exact_kernel(const double *__restrict__ k, const double *__restrict__ alpha,
                                                       const double *__restrict__ p, int64_t n, double *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = dig_nb::nb_exact(k[i], alpha[i], p[i]);
}
}  // namespace

extern "C" int dig_nb_pvalue_exact(const double *k_d, const double *alpha_d, const double *p_d, int64_t n,
                                   double *pval_out_d, void *stream)
{
    DIG_CHECK_ARG(n >= 0, "negative size");
    if (n == 0) return DIG_OK;
    DIG_CHECK_ARG(k_d && alpha_d && p_d && pval_out_d, "null pointer");
    int64_t blocks = (n + 127) / 128;
    const int64_t cap = (int64_t)dig::sm_count() * 32;
    if (blocks > cap) blocks = cap;
    nb_exact_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(k_d, alpha_d, p_d, n, pval_out_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}
