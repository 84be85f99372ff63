// Gene-level burden stage fused on the device: scale-factor sums, then all 13 NB tests + Fisher per gene.
//
// Replaces the table arithmetic of run_gene_model (transfer_tools.py:809-861): load_pretrained_model's
// ALPHA/THETA (:17-19, nb_model.py:237-241), Pi_TRUNC/Pi_NONSYN (genic_driver_tools.py:123, transfer_tools.py:34),
// the synonymous scale factor cj (:813-815), gene_expected_muts_nb (:331-341), gene_pvalue_burden_nb (:394-456),
// gene_pvalue_burden_nb_by_sample (:484-592), gene_pvalue_indel (:709-729) and the Fisher combine (:860-861).
#include <cooperative_groups.h>

#include "nb_math.cuh"

namespace {

using namespace dig_nb;

// sums[0] = sum_{g != TP53} MU * Pi_SYN          sums[1] = sum_{g not CGC} Pi_INDEL * ALPHA * THETA
// sums[2] = sum_{g not CGC} OBS_INDEL
// Fixed summation order whatever the timing: ONE cluster of eight blocks (8192 threads, so 20 k genes are two or three
// dependent loads per thread instead of twenty), strided partial sums, a tree per block, then block 0 adds the eight
// block sums in rank order through distributed shared memory.  No global scratch, no atomics.
constexpr int SUMS_CLUSTER = 8;
__global__ void __cluster_dims__(SUMS_CLUSTER, 1, 1) __launch_bounds__(1024) gene_scale_sums_kernel(
    const double *__restrict__ mu, const double *__restrict__ sigma, const double *__restrict__ P,
    const double *__restrict__ pi_indel, const int64_t *__restrict__ obs, const uint8_t *__restrict__ cgc,
    int64_t tp53, int64_t n, double *__restrict__ sums)
{
    __shared__ double sh[3][1024];
    cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
    const unsigned int rank = cluster.block_rank();
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    // unrolled so that the (independent) loads of four strides are in flight together; the additions keep their order
#pragma unroll 4
    for (int64_t g = (int64_t)rank * 1024 + threadIdx.x; g < n; g += (int64_t)SUMS_CLUSTER * 1024) {
        const double m = mu[g], s = sigma[g];
        // pandas' Series.sum() skips NaN (skipna=True): a gene whose windows hold no countable context (P = NaN)
        // must not poison the cohort-wide scale factors (transfer_tools.py:814, :699-700)
        const double t0 = m * P[g * 4 + 0];
        if (g != tp53 && !isnan(t0)) a0 += t0;
        if (cgc == nullptr || !cgc[g]) {
            const double alpha = (m * m) / (s * s);
            const double theta = (s * s) / m;
            const double t1 = pi_indel[g] * alpha * theta;
            if (!isnan(t1)) a1 += t1;
            a2 += (double)obs[g * 5 + 4];
        }
    }
    sh[0][threadIdx.x] = a0;
    sh[1][threadIdx.x] = a1;
    sh[2][threadIdx.x] = a2;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
            sh[2][threadIdx.x] += sh[2][threadIdx.x + o];
        }
        __syncthreads();
    }
    cluster.sync();                                        // every block's three sums are in its sh[.][0]
    if (rank == 0 && threadIdx.x < 3) {
        double acc = 0.0;
        for (unsigned int r = 0; r < (unsigned int)SUMS_CLUSTER; ++r) acc += *cluster.map_shared_rank(&sh[threadIdx.x][0], r);
        sums[threadIdx.x] = acc;
    }
    cluster.sync();                                        // nobody leaves while block 0 is still reading its shared memory
}

// test t of gene g: t in 0..5 count burden (SYN, MIS, NONS, SPL, TRUNC, NONSYN), 6..11 sample burden, 12 indel
__global__ void __launch_bounds__(128) gene_test_kernel(
    const double *__restrict__ mu, const double *__restrict__ sigma, const double *__restrict__ P,
    const double *__restrict__ pi_indel, const int64_t *__restrict__ obs, const int64_t *__restrict__ nsamp,
    int64_t n, const double *__restrict__ sums, double n_syn,
    double scale_factor, double *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double cj = isnan(scale_factor) ? (isnan(n_syn) ? sums[3] : n_syn) / sums[0] : scale_factor;
    const double t_indel = sums[2] / sums[1];
    // gene fastest: the 32 lanes of a warp run the SAME test for 32 consecutive genes, so the branches on t are uniform,
    // the loads and the stores are contiguous and the continued fractions of a warp have similar lengths
    // (test fastest, 13 tests of two or three genes per warp: 62 us for 260 k p-values)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * 13; i += stride) {
        const int t = (int)(i / n);
        const int64_t g = i - (int64_t)t * n;
        const double m = mu[g], s = sigma[g];
        const double alpha = (m * m) / (s * s);
        const double theta0 = (s * s) / m;
        const double p_syn = P[g * 4 + 0], p_mis = P[g * 4 + 1], p_nons = P[g * 4 + 2], p_spl = P[g * 4 + 3];
        const double p_trunc = p_nons + p_spl;
        const double p_nonsyn = p_mis + p_trunc;
        double pi, k, theta;
        if (t == 12) {
            pi = pi_indel[g];
            k = (double)obs[g * 5 + 4];
            theta = theta0 * t_indel;
        } else {
            const int c = t % 6;
            pi = c == 0 ? p_syn : c == 1 ? p_mis : c == 2 ? p_nons : c == 3 ? p_spl : c == 4 ? p_trunc : p_nonsyn;
            theta = theta0 * cj;
            if (t < 6) {
                const double o0 = (double)obs[g * 5 + 0], o1 = (double)obs[g * 5 + 1], o2 = (double)obs[g * 5 + 2],
                             o3 = (double)obs[g * 5 + 3];
                k = c == 0 ? o0 : c == 1 ? o1 : c == 2 ? o2 : c == 3 ? o3 : c == 4 ? o2 + o3 : o1 + (o2 + o3);
            } else {
                k = (double)nsamp[g * 7 + (c < 4 ? c : c)];      // nsamp columns: SYN MIS NONS SPL TRUNC NONSYN INDEL
            }
        }
        const double expv = __dmul_rn(__dmul_rn(alpha, theta), pi);
        const double p = __ddiv_rn(1.0, __dadd_rn(__dmul_rn(theta, pi), 1.0));
        const double pval = nb_midp(k, alpha, p);
        // rows: 0-5 EXP_*, 6-11 PVAL_*_BURDEN, 12-17 PVAL_*_BURDEN_SAMPLE, 18 EXP_INDEL, 19 PVAL_INDEL_BURDEN
        if (t < 6) {
            out[(int64_t)t * n + g] = expv;
            out[(int64_t)(6 + t) * n + g] = pval;
        } else if (t < 12) {
            out[(int64_t)(6 + t) * n + g] = pval;
        } else {
            out[18 * n + g] = expv;
            out[19 * n + g] = pval;
        }
        if (t == 0) {
            out[21 * n + g] = alpha;
            out[22 * n + g] = theta;
            out[23 * n + g] = theta0 * t_indel;
            out[24 * n + g] = pi_indel[g];
            out[25 * n + g] = p_trunc;
            out[26 * n + g] = p_nonsyn;
        }
    }
}

__global__ void __launch_bounds__(256) gene_fisher_kernel(int64_t n, double *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += stride)
        out[20 * n + g] = fisher2(out[(6 + 4) * n + g], out[19 * n + g]);     // PVAL_TRUNC_BURDEN x PVAL_INDEL_BURDEN
}

__global__ void __launch_bounds__(256) seq_freq_kernel(const unsigned long long *__restrict__ counts,
                                                       const unsigned long long *__restrict__ totals, int n_sub,
                                                       double *__restrict__ freq)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_sub) freq[j] = (double)counts[j] / (double)totals[j / 3];
}

__global__ void __launch_bounds__(256) size_ratio_kernel(const int64_t *__restrict__ num, const int64_t *__restrict__ den, int64_t n,
                                                         double *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __ddiv_rn((double)num[i], (double)den[i]);
}

}  // namespace

extern "C" {

int dig_size_ratio(const int64_t *num_d, const int64_t *den_d, int64_t n, double *out_d, void *stream)
{
    DIG_CHECK_ARG(n >= 0, "negative size");
    if (n == 0) return DIG_OK;
    DIG_CHECK_ARG(num_d && den_d && out_d, "null pointer");
    size_ratio_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(num_d, den_d, n, out_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

int dig_sequence_freq(const unsigned long long *subst_counts_d, const unsigned long long *ctx_totals_d, int n_ctx,
                      double *freq_d, void *stream)
{
    DIG_CHECK_ARG(n_ctx > 0 && subst_counts_d && ctx_totals_d && freq_d, "bad arguments");
    const int n_sub = 3 * n_ctx;
    seq_freq_kernel<<<(n_sub + 255) / 256, 256, 0, (cudaStream_t)stream>>>(subst_counts_d, ctx_totals_d, n_sub, freq_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

int dig_gene_scale_sums(const double *mu_d, const double *sigma_d, const double *p_d, const double *pi_indel_d,
                        const int64_t *obs_d, const uint8_t *cgc_mask_d, int64_t tp53, int64_t n_gene,
                        double *sums_d, void *stream)
{
    DIG_CHECK_ARG(n_gene >= 0 && sums_d, "bad arguments");
    DIG_CHECK_ARG(n_gene == 0 || (mu_d && sigma_d && p_d && pi_indel_d && obs_d), "null pointer");
    gene_scale_sums_kernel<<<SUMS_CLUSTER, 1024, 0, (cudaStream_t)stream>>>(mu_d, sigma_d, p_d, pi_indel_d, obs_d, cgc_mask_d,
                                                                 tp53, n_gene, sums_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

int dig_gene_burden_test(const double *mu_d, const double *sigma_d, const double *p_d, const double *pi_indel_d,
                         const int64_t *obs_d, const int64_t *nsamp_d, int64_t n_gene, const double *sums_d,
                         double n_syn, double scale_factor, double *out_d, void *stream)
{
    DIG_CHECK_ARG(n_gene >= 0, "negative size");
    if (n_gene == 0) return DIG_OK;
    DIG_CHECK_ARG(mu_d && sigma_d && p_d && pi_indel_d && obs_d && nsamp_d && sums_d && out_d, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (n_gene * 13 + 127) / 128;
    const int64_t cap = (int64_t)dig::sm_count() * 32;
    if (blocks > cap) blocks = cap;
    gene_test_kernel<<<(unsigned)blocks, 128, 0, st>>>(mu_d, sigma_d, p_d, pi_indel_d, obs_d, nsamp_d, n_gene, sums_d,
                                                       n_syn, scale_factor, out_d);
    DIG_CHECK_LAUNCH();
    gene_fisher_kernel<<<(unsigned)((n_gene + 255) / 256), 256, 0, st>>>(n_gene, out_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

}

// ---------------------------------------------------------------------------------------
// Secondary gene tests: dN/dS-corrected expectations and burden p-values and the NB likelihood-
// ratio selection tests, one thread per gene.
//   gene_expected_muts_dnds  transfer_tools.py:363-392   EXP_x, T_SYN (_mle_t :1263-1271), MRFOLD (:1273-1276), EXP_x_ML
//   gene_pvalue_burden_dnds  :617-653                    nb_midp(OBS_x, ALPHA, 1 / (EXP_x_ML / ALPHA + 1))
//   gene_pvalue_sel_nb       :655-676 + _llr_test_nb :1172-1214   chi2.sf(-2 (ll0 - ll_k), df = 1 | 2)
// Arithmetic follows the reference's operation order without FMA contraction.
// ---------------------------------------------------------------------------------------
namespace {

using namespace dig_nb;

// scipy.stats.nbinom.logpmf(k, alpha, 1 / (1 + theta))  (_ll_nb, transfer_tools.py:1253-1255)
__device__ inline double ll_nb(double k, double alpha, double theta)
{
    const double p = __ddiv_rn(1.0, __dadd_rn(1.0, theta));
    const double q = 1.0 - p;
    if (isnan(p) || isnan(k) || isnan(alpha)) return __longlong_as_double(0x7ff8000000000000LL);
    if (q <= 0.0) return k == 0.0 ? 0.0 : -INFINITY;        // all mass at 0
    if (p <= 0.0) return -INFINITY;
    return log_nb_density(k, alpha, p, q);
}

__global__ void __launch_bounds__(128) gene_dnds_sel_kernel(const double *__restrict__ alpha,
                                                            const double *__restrict__ theta,
                                                            const double *__restrict__ pi,    // [n, 6]
                                                            const double *__restrict__ obs,   // [n, 6]
                                                            int64_t n, double *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += stride) {
        const double a = alpha[g], t = theta[g];
        double P[6], O[6], E[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            P[j] = pi[g * 6 + j];
            O[j] = obs[g * 6 + j];
            E[j] = __dmul_rn(__dmul_rn(a, t), P[j]);
            out[(int64_t)j * n + g] = E[j];
        }
        // _mle_t(OBS_SYN, 1, ALPHA, THETA * Pi_SYN)
        const double th_syn = __dmul_rn(t, P[0]);
        double tml = __ddiv_rn(__dadd_rn(__dadd_rn(O[0], a), -1.0), __dadd_rn(1.0, __ddiv_rn(1.0, th_syn)));
        if (a <= 1.0) {
            const double lo = __dmul_rn(a, th_syn);
            tml = tml > lo ? tml : lo;                       // Python max(lo, tml): the first argument wins ties and NaN
        }
        const double ratio = __ddiv_rn(tml, E[0]);
        const double mrf = ratio > 1e-10 ? ratio : 1e-10;    // Python max(1e-10, ratio)
        out[6 * n + g] = tml;
        out[7 * n + g] = mrf;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const double eml = __dmul_rn(E[j], mrf);
            out[(int64_t)(8 + j) * n + g] = eml;
            const double p = __ddiv_rn(1.0, __dadd_rn(__ddiv_rn(eml, a), 1.0));
            out[(int64_t)(14 + j) * n + g] = nb_midp(O[j], a, p);
        }
        // likelihood-ratio selection tests on SYN (0), MIS (1), TRUNC (4)
        const int cls[3] = {0, 1, 4};
        double l0[3], l1[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int j = cls[c];
            l0[c] = ll_nb(O[j], a, __dmul_rn(__dmul_rn(t, P[j]), mrf));
            l1[c] = ll_nb(O[j], a, __ddiv_rn(O[j], a));
        }
        const double ll0 = __dadd_rn(__dadd_rn(l0[0], l0[1]), l0[2]);
        const double ll_syn = __dadd_rn(__dadd_rn(l1[0], l0[1]), l0[2]);
        const double ll_mis = __dadd_rn(__dadd_rn(l0[0], l1[1]), l0[2]);
        const double ll_tr = __dadd_rn(__dadd_rn(l0[0], l0[1]), l1[2]);
        const double ll_ns = __dadd_rn(__dadd_rn(l0[0], l1[1]), l1[2]);
        out[20 * n + g] = chi2_sf(-2.0 * (ll0 - ll_syn), 1);
        out[21 * n + g] = chi2_sf(-2.0 * (ll0 - ll_mis), 1);
        out[22 * n + g] = chi2_sf(-2.0 * (ll0 - ll_tr), 1);
        out[23 * n + g] = chi2_sf(-2.0 * (ll0 - ll_ns), 2);
    }
}

// selection_coefficient (transfer_tools.py:1279-1292): SEL = (OBS + 1e-16) / (EXP + 1e-16) and the LLR p-value of
// nbinom(ALPHA, 1/(1 + THETA*Pi)) against nbinom(ALPHA, 1/(1 + THETA*Pi*SEL))
__global__ void __launch_bounds__(128) selection_coef_kernel(const double *__restrict__ obs, const double *__restrict__ ex,
                                                             const double *__restrict__ alpha,
                                                             const double *__restrict__ theta,
                                                             const double *__restrict__ pi, int64_t n,
                                                             double *__restrict__ sel, double *__restrict__ pval)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += stride) {
        const double s = __ddiv_rn(__dadd_rn(obs[g], 1e-16), __dadd_rn(ex[g], 1e-16));
        sel[g] = s;
        if (pval) {
            const double tp = __dmul_rn(theta[g], pi[g]);
            const double ll0 = ll_nb(obs[g], alpha[g], tp);
            const double ll1 = ll_nb(obs[g], alpha[g], __dmul_rn(tp, s));
            pval[g] = chi2_sf(-2.0 * (ll0 - ll1), 1);
        }
    }
}

}  // namespace

extern "C" int dig_gene_dnds_sel(const double *alpha_d, const double *theta_d, const double *pi6_d, const double *obs6_d,
                                 int64_t n_gene, double *out_d, void *stream)
{
    DIG_CHECK_ARG(n_gene >= 0, "negative size");
    if (n_gene == 0) return DIG_OK;
    DIG_CHECK_ARG(alpha_d && theta_d && pi6_d && obs6_d && out_d, "null pointer");
    int64_t blocks = (n_gene + 127) / 128;
    gene_dnds_sel_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(alpha_d, theta_d, pi6_d, obs6_d, n_gene,
                                                                            out_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

extern "C" int dig_selection_coefficient(const double *obs_d, const double *exp_d, const double *alpha_d,
                                         const double *theta_d, const double *pi_d, int64_t n, double *sel_d,
                                         double *pval_d, void *stream)
{
    DIG_CHECK_ARG(n >= 0, "negative size");
    if (n == 0) return DIG_OK;
    DIG_CHECK_ARG(obs_d && exp_d && sel_d, "null pointer");
    DIG_CHECK_ARG(pval_d == nullptr || (alpha_d && theta_d && pi_d), "p-values need alpha, theta and pi");
    int64_t blocks = (n + 127) / 128;
    selection_coef_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(obs_d, exp_d, alpha_d, theta_d, pi_d, n,
                                                                             sel_d, pval_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

// Row-level likelihood-ratio selection tests with the caller's MRFOLD (and T_SYN): _llr_test_nb
// (transfer_tools.py:1172-1213, classes SYN / MIS / TRUNC) and _llr_test_gamma_poiss (:1215-1252, classes SYN / MIS /
// NONS plus the gamma prior term of T_SYN, which is common to all five likelihoods but is added in the reference's
// order so that NaN / inf propagate the same way).  One thread per row; out [4, n] = p_syn, p_mis, p_third, p_nonsyn.
namespace {

__global__ void __launch_bounds__(128) gene_llr_kernel(int model, const double *__restrict__ alpha,
                                                       const double *__restrict__ theta,
                                                       const double *__restrict__ pi,     // [n, 3]
                                                       const double *__restrict__ obs,    // [n, 3]
                                                       const double *__restrict__ mrfold, const double *__restrict__ t_syn,
                                                       int64_t n, double *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += stride) {
        double r[4];
        llr_row(model, alpha[g], theta[g], pi + g * 3, obs + g * 3, mrfold[g], t_syn ? t_syn[g] : 0.0, r);
#pragma unroll
        for (int j = 0; j < 4; ++j) out[(int64_t)j * n + g] = r[j];
    }
}

}  // namespace

extern "C" int dig_gene_llr_test(int model, const double *alpha_d, const double *theta_d, const double *pi3_d,
                                 const double *obs3_d, const double *mrfold_d, const double *t_syn_d, int64_t n,
                                 double *out_d, void *stream)
{
    DIG_CHECK_ARG(n >= 0, "negative size");
    DIG_CHECK_ARG(model == DIG_LLR_NB || model == DIG_LLR_GAMMA_POISSON, "unknown model");
    if (n == 0) return DIG_OK;
    DIG_CHECK_ARG(alpha_d && theta_d && pi3_d && obs3_d && mrfold_d && out_d, "null pointer");
    DIG_CHECK_ARG(model == DIG_LLR_NB || t_syn_d, "the gamma-Poisson model needs T_SYN");
    int64_t blocks = (n + 127) / 128;
    gene_llr_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(model, alpha_d, theta_d, pi3_d, obs3_d, mrfold_d,
                                                                       t_syn_d, n, out_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}
