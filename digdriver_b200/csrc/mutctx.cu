// K3: mutation context lookup with REF check (mutation_contexts_by_chrom, sequence_tools.py:130-178).
// One thread per mutation row (latency bound, tiny data).
#include "dig_common.cuh"

namespace {

// The k-mer (at most 15 bases) lies inside two consecutive packed words and its N bits inside two consecutive mask
// words: four independent loads give the context AND the base under the mutation for the REF check, so a mutation
// costs one memory round trip instead of a chain of dependent 4-byte reads.
__global__ void __launch_bounds__(256) mutctx_kernel(
    const uint32_t *__restrict__ p2, const uint32_t *__restrict__ nmask, int64_t last_p, int64_t last_m,
    const int64_t *__restrict__ chrom_off, const int64_t *__restrict__ chrom_len,
    const int32_t *__restrict__ mut_chrom, const int64_t *__restrict__ mut_start,
    const uint8_t *__restrict__ mut_ref, int64_t n_mut, int n_up, int n_down, int32_t *__restrict__ ctx_out)
{
    const int klen = n_up + n_down + 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_mut; i += stride) {
        const int32_t c = mut_chrom[i];
        const int64_t p = mut_start[i];
        const int64_t L = __ldg(chrom_len + c);
        const int64_t off = __ldg(chrom_off + c);
        int32_t ctx = -1;
        // a context that would leave the chromosome is dropped whatever the REF check says
        if (p - n_up >= 0 && p + n_down < L) {
            const int64_t f = off + p - n_up;                          // first base of the k-mer
            const int64_t w = f >> 4, m = f >> 5;
            const unsigned long long pk = ((unsigned long long)__ldg(p2 + w) << 32) | __ldg(p2 + (w < last_p ? w + 1 : last_p));
            const unsigned long long mk = ((unsigned long long)__ldg(nmask + m) << 32) | __ldg(nmask + (m < last_m ? m + 1 : last_m));
            const uint32_t key = (uint32_t)(pk >> (64 - 2 * ((int)(f & 15) + klen))) & ((1u << (2 * klen)) - 1u);
            const uint32_t bad = (uint32_t)(mk >> (64 - ((int)(f & 31) + klen))) & ((1u << klen) - 1u);
            const uint32_t base = (key >> (2 * n_down)) & 3u;          // the base under the mutation ...
            const bool base_n = ((bad >> n_down) & 1u) != 0u;          // ... and whether it is not A/C/G/T
            auto mismatch = [&](int64_t row) -> bool {
                const uint32_t ref = mut_ref[row];
                return ref > 3u || base_n || base != ref;              // an N in the genome never equals A/C/G/T
            };
            bool dropped = mismatch(i);
            // same-START run: the reference re-uses the previous row's context string, so a REF
            // mismatch earlier in the run drops every later row of the run (sequence_tools.py:150-151)
            for (int64_t j = i - 1; !dropped && j >= 0 && mut_chrom[j] == c && mut_start[j] == p; --j)
                dropped = mismatch(j);
            if (!dropped && bad == 0u) ctx = (int32_t)key;
        }
        ctx_out[i] = ctx;
    }
}

// 3K-bin histogram of substitutions: bin = 3 * ctx + rank(ALT among the non-REF bases, alphabetical),
// i.e. the sorted 'CTX>CTX2' order of mk_trans_idx (sequence_tools.py:282-289) for trinucleotides.
__global__ void __launch_bounds__(256) subst_hist_kernel(const int32_t *__restrict__ ctx,
                                                         const uint8_t *__restrict__ alt, int64_t n, int n_down,
                                                         int n_bins, unsigned long long *__restrict__ counts)
{
    extern __shared__ unsigned int sh[];
    for (int k = threadIdx.x; k < n_bins; k += blockDim.x) sh[k] = 0u;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int32_t c = ctx[i];
        const uint32_t a = alt[i];
        if (c < 0 || a > 3u) continue;
        const uint32_t ref = ((uint32_t)c >> (2 * n_down)) & 3u;
        if (a == ref) continue;
        atomicAdd(sh + 3 * c + (int)(a > ref ? a - 1u : a), 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n_bins; k += blockDim.x)
        if (sh[k]) atomicAdd(counts + k, (unsigned long long)sh[k]);
}

}  // namespace

extern "C" int dig_substitution_counts(const int32_t *ctx_d, const uint8_t *alt_d, int64_t n, int n_up, int n_down,
                                       unsigned long long *counts_d, void *stream)
{
    DIG_CHECK_ARG(n >= 0, "negative size");
    DIG_CHECK_ARG(n_up >= 0 && n_down >= 0 && n_up + n_down <= 5, "unsupported context size");
    DIG_CHECK_ARG(counts_d != nullptr, "null pointer");
    const int n_bins = 3 << (2 * (n_up + n_down + 1));
    cudaStream_t st = (cudaStream_t)stream;
    DIG_CUDA(cudaMemsetAsync(counts_d, 0, (size_t)n_bins * sizeof(unsigned long long), st));
    if (n == 0) return DIG_OK;
    DIG_CHECK_ARG(ctx_d && alt_d, "null pointer");
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)dig::sm_count() * 4;
    if (blocks > cap) blocks = cap;
    subst_hist_kernel<<<(unsigned)blocks, 256, (size_t)n_bins * sizeof(unsigned int), st>>>(ctx_d, alt_d, n, n_down,
                                                                                          n_bins, counts_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

extern "C" int dig_mutation_contexts(const uint32_t *packed2_d, const uint32_t *nmask_d, int64_t n_bases,
                                     const int64_t *chrom_off_d, const int64_t *chrom_len_d,
                                     const int32_t *mut_chrom_d, const int64_t *mut_start_d,
                                     const uint8_t *mut_ref_d, int64_t n_mut, int n_up, int n_down,
                                     int32_t *ctx_out_d, void *stream)
{
    DIG_CHECK_ARG(n_mut >= 0 && n_bases >= 0, "negative size");
    DIG_CHECK_ARG(n_up >= 0 && n_down >= 0 && n_up + n_down <= 14, "unsupported context size");
    if (n_mut == 0) return DIG_OK;
    DIG_CHECK_ARG(packed2_d && nmask_d && chrom_off_d && chrom_len_d && mut_chrom_d && mut_start_d && mut_ref_d &&
                      ctx_out_d,
                  "null pointer");
    int64_t blocks = (n_mut + 255) / 256;
    const int64_t cap = (int64_t)dig::sm_count() * 32;
    if (blocks > cap) blocks = cap;
    mutctx_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        packed2_d, nmask_d, (((n_bases + 31) >> 5) << 1) - 1, ((n_bases + 31) >> 5) - 1, chrom_off_d, chrom_len_d,
        mut_chrom_d, mut_start_d, mut_ref_d, n_mut, n_up, n_down, ctx_out_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}
