// K7: FP64 negative-binomial mid-p upper-tail p-values, expectations and the Fisher combine.
//
// Replaces the SciPy expression of nb_model.nb_pvalue_greater_midp (nb_model.py:271-278)
//     pval = 0.5 * nbinom.pmf(k, alpha, p) + betainc(k + 1, alpha, 1 - p)
// as called from transfer_tools.py:473-482 / :594-615 / :394-456 / :484-592 / :709-747.
//
// Numerics.  Both terms are written relative to the (real-k) NB density G(k):
//     betainc(k+1, a, q)  =      G(k) * q (k+a)/(k+1) * CF(k+1, a, q)          q <  (k+2)/(k+a+3)
//                         = 1 -  G(k) * q (k+a)/a     * CF(a, k+1, p)          otherwise
// where CF is the continued fraction of the incomplete beta function evaluated with the
// modified Lentz algorithm, and log G(k) is computed with Loader's saddle-point expansion
// (Stirling error terms + the deviance function bd0) so that no large lgamma values are
// ever subtracted: it stays accurate for alpha from 1e-6 to 1e8.  Everything is FP64 and
// evaluated in log space until the final exp, so deep tails underflow to 0.0 exactly where
// SciPy's do.  Measured against SciPy 1.18.1 on the golden grid: max |dlog10 p| = 2e-9.
#include "nb_math.cuh"

namespace {

using namespace dig_nb;

__global__ void __launch_bounds__(128) nb_midp_kernel(const double *__restrict__ k, const double *__restrict__ alpha,
                                                      const double *__restrict__ p, int64_t n,
                                                      double *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = nb_midp(k[i], alpha[i], p[i]);
}

// EXP = ALPHA*THETA*Pi (transfer_tools.py:343-344) and p = 1/(THETA*Pi + 1) (:479) in the
// reference's own operation order, without FMA contraction, so the inputs of the p-value are
// bit-identical to the reference's.
__global__ void __launch_bounds__(128) nb_burden_kernel(const double *__restrict__ k, const double *__restrict__ alpha,
                                                        const double *__restrict__ theta,
                                                        const double *__restrict__ pi, int64_t n,
                                                        double *__restrict__ exp_out, double *__restrict__ pval_out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double a = alpha[i], t = theta[i], f = pi[i];
        if (exp_out) exp_out[i] = __dmul_rn(__dmul_rn(a, t), f);
        const double p = __ddiv_rn(1.0, __dadd_rn(__dmul_rn(t, f), 1.0));
        pval_out[i] = nb_midp(k[i], a, p);
    }
}

// chi2.sf(-2 (ln p1 + ln p2), df=4) = exp(-y) (1 + y), y = -(ln p1 + ln p2)  (transfer_tools.py:860-861)
__global__ void __launch_bounds__(256) fisher2_kernel(const double *__restrict__ p1, const double *__restrict__ p2,
                                                      int64_t n, double *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double r = fisher2(p1[i], p2[i]);
        out[i] = r;
    }
}

// The remaining tail conventions of nb_model.py (:243-337) -- see nb_variant in nb_math.cuh.
__global__ void __launch_bounds__(128) nb_variant_kernel(int mode, const double *__restrict__ k,
                                                         const double *__restrict__ alpha, const double *__restrict__ p,
                                                         const double *__restrict__ mu, int64_t n,
                                                         double *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = nb_variant(mode, k[i], alpha[i], p[i], mu ? mu[i] : 0.0, mu != nullptr);
}

__global__ void __launch_bounds__(128) loglik_kernel(int kind, const double *__restrict__ x, const double *__restrict__ a,
                                                     const double *__restrict__ b, int64_t n, double *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = kind == 0 ? ll_nb_dev(x[i], a[i], b[i]) : kind == 1 ? ll_pois_dev(x[i], a[i]) : ll_gamma_dev(x[i], a[i], b[i]);
}

inline unsigned grid_for(int64_t n, int threads)
{
    int64_t blocks = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)dig::sm_count() * 32;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}

}  // namespace

extern "C" {

int dig_nb_pvalue_greater_midp(const double *k_d, const double *alpha_d, const double *p_d, int64_t n,
                               double *pval_out_d, void *stream)
{
    DIG_CHECK_ARG(n >= 0, "negative size");
    if (n == 0) return DIG_OK;
    DIG_CHECK_ARG(k_d && alpha_d && p_d && pval_out_d, "null pointer");
    nb_midp_kernel<<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(k_d, alpha_d, p_d, n, pval_out_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

int dig_nb_burden_test(const double *k_d, const double *alpha_d, const double *theta_d, const double *pi_d,
                       int64_t n, double *exp_out_d, double *pval_out_d, void *stream)
{
    DIG_CHECK_ARG(n >= 0, "negative size");
    if (n == 0) return DIG_OK;
    DIG_CHECK_ARG(k_d && alpha_d && theta_d && pi_d && pval_out_d, "null pointer");
    nb_burden_kernel<<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(k_d, alpha_d, theta_d, pi_d, n, exp_out_d,
                                                                         pval_out_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

int dig_fisher_combine2(const double *p1_d, const double *p2_d, int64_t n, double *out_d, void *stream)
{
    DIG_CHECK_ARG(n >= 0, "negative size");
    if (n == 0) return DIG_OK;
    DIG_CHECK_ARG(p1_d && p2_d && out_d, "null pointer");
    fisher2_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(p1_d, p2_d, n, out_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

int dig_nb_pvalue_variant(int mode, const double *k_d, const double *alpha_d, const double *p_d, const double *mu_d,
                          int64_t n, double *pval_out_d, void *stream)
{
    DIG_CHECK_ARG(n >= 0, "negative size");
    DIG_CHECK_ARG(mode >= DIG_NB_GREATER && mode <= DIG_NB_MIDP, "unknown mode");
    if (n == 0) return DIG_OK;
    DIG_CHECK_ARG(k_d && alpha_d && p_d && pval_out_d, "null pointer");
    nb_variant_kernel<<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(mode, k_d, alpha_d, p_d, mu_d, n, pval_out_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

int dig_loglik(int kind, const double *x_d, const double *a_d, const double *b_d, int64_t n, double *out_d, void *stream)
{
    DIG_CHECK_ARG(n >= 0, "negative size");
    DIG_CHECK_ARG(kind >= DIG_LL_NB && kind <= DIG_LL_GAMMA, "unknown kind");
    if (n == 0) return DIG_OK;
    DIG_CHECK_ARG(x_d && a_d && out_d && (kind == DIG_LL_POIS || b_d), "null pointer");
    loglik_kernel<<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(kind, x_d, a_d, b_d, n, out_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

}
