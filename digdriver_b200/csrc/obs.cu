// K5: observed mutation counts per element / per gene.
//
// Elements: interval stabbing of sorted bed6 blocks (binary search + backward walk bounded by a
// prefix-max of block ends) replaces `bedtools intersect -wa -wb`; an (element, sample) open-
// addressing hash table in HBM replaces the pandas group-bys.  All accumulations are integer
// atomics, so results are deterministic and bit-exact.
#include "dig_common.cuh"

namespace {

__device__ __forceinline__ int64_t upper_bound_ge(const int64_t *__restrict__ a, int64_t n, int64_t v)
{
    // first index with a[idx] >= v
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// Calls f(elt) once for every distinct element that mutation [ks, ke) overlaps.
template <typename F>
__device__ __forceinline__ void for_each_hit(const int64_t *__restrict__ bks, const int64_t *__restrict__ bke,
                                             const int64_t *__restrict__ pmax, const int32_t *__restrict__ belt,
                                             int64_t n_blk, int64_t ks, int64_t ke, F f)
{
    if (ke <= ks) return;
    const int64_t hi = upper_bound_ge(bks, n_blk, ke);       // blocks [0, hi) start before the mutation ends
    for (int64_t idx = hi - 1; idx >= 0; --idx) {
        if (__ldg(pmax + idx) <= ks) break;                  // nothing at or before idx reaches the mutation
        if (__ldg(bke + idx) <= ks) continue;
        const int32_t e = __ldg(belt + idx);
        bool dup = false;                                    // an already visited block of the same element?
        for (int64_t j = hi - 1; j > idx && !dup; --j) dup = (__ldg(belt + j) == e) && (__ldg(bke + j) > ks);
        if (!dup) f(e);
    }
}

__global__ void __launch_bounds__(256) count_hits_kernel(
    const int64_t *__restrict__ bks, const int64_t *__restrict__ bke, const int64_t *__restrict__ pmax,
    const int32_t *__restrict__ belt, int64_t n_blk, const int64_t *__restrict__ mks,
    const int64_t *__restrict__ mke, int64_t n_mut, unsigned long long *n_hits)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long local = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_mut; i += stride)
        for_each_hit(bks, bke, pmax, belt, n_blk, mks[i], mke[i], [&](int32_t) { ++local; });
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(n_hits, local);
}

// find-or-insert; returns slot or -1 on overflow
__device__ __forceinline__ int64_t table_slot(unsigned long long *keys, int64_t capacity, unsigned long long key)
{
    const int64_t mask = capacity - 1;
    int64_t slot = (int64_t)(dig::mix64(key) & (unsigned long long)mask);
    for (int64_t probes = 0; probes < capacity; ++probes) {
        const unsigned long long prev = atomicCAS(keys + slot, 0ull, key);
        if (prev == 0ull || prev == key) return slot;
        slot = (slot + 1) & mask;
    }
    return -1;
}

__global__ void __launch_bounds__(256) elt_insert_kernel(
    const int64_t *__restrict__ bks, const int64_t *__restrict__ bke, const int64_t *__restrict__ pmax,
    const int32_t *__restrict__ belt, int64_t n_blk, const int64_t *__restrict__ mks,
    const int64_t *__restrict__ mke, const int32_t *__restrict__ msample, const uint8_t *__restrict__ mindel,
    int64_t n_mut, unsigned long long *keys, uint32_t *snv, uint32_t *indel, int64_t capacity, int32_t *status)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_mut; i += stride) {
        const unsigned long long sample = (unsigned long long)(uint32_t)msample[i];
        const bool is_indel = mindel[i] != 0;
        for_each_hit(bks, bke, pmax, belt, n_blk, mks[i], mke[i], [&](int32_t e) {
            const unsigned long long key = (((unsigned long long)(uint32_t)e << 32) | sample) + 1ull;
            const int64_t slot = table_slot(keys, capacity, key);
            if (slot < 0) {
                *status = 1;
                return;
            }
            atomicAdd((is_indel ? indel : snv) + slot, 1u);
        });
    }
}

__global__ void __launch_bounds__(256) elt_sample_totals_kernel(const unsigned long long *__restrict__ keys,
                                                                const uint32_t *__restrict__ snv,
                                                                const uint32_t *__restrict__ indel, int64_t capacity,
                                                                int rows_mode, unsigned long long *sample_tot)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < capacity; s += stride) {
        const unsigned long long key = keys[s];
        if (key == 0ull) continue;
        const uint32_t sample = (uint32_t)((key - 1ull) & 0xFFFFFFFFull);
        // rows_mode: one per (element, sample) row = df.SAMPLE.value_counts() of the per-sample-per-element table
        atomicAdd(sample_tot + sample, rows_mode ? 1ull : (unsigned long long)snv[s] + (unsigned long long)indel[s]);
    }
}

__global__ void __launch_bounds__(256) elt_finalize_kernel(const unsigned long long *__restrict__ keys,
                                                           const uint32_t *__restrict__ snv,
                                                           const uint32_t *__restrict__ indel, int64_t capacity,
                                                           const unsigned long long *__restrict__ sample_tot,
                                                           int64_t max_per_sample, int64_t max_per_elt_sample,
                                                           int64_t n_elt, int64_t *obs)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < capacity; s += stride) {
        const unsigned long long key = keys[s];
        if (key == 0ull) continue;
        const uint32_t sample = (uint32_t)((key - 1ull) & 0xFFFFFFFFull);
        const int64_t e = (int64_t)((key - 1ull) >> 32);
        if (e >= n_elt) continue;
        if ((int64_t)sample_tot[sample] > max_per_sample) continue;      // hypermutator black-list
        int64_t a = snv[s], b = indel[s];
        if (a > max_per_elt_sample) a = max_per_elt_sample;
        if (b > max_per_elt_sample) b = max_per_elt_sample;
        atomicAdd(reinterpret_cast<unsigned long long *>(obs + 3 * e + 0), 1ull);
        if (a) atomicAdd(reinterpret_cast<unsigned long long *>(obs + 3 * e + 1), (unsigned long long)a);
        if (b) atomicAdd(reinterpret_cast<unsigned long long *>(obs + 3 * e + 2), (unsigned long long)b);
    }
}

__global__ void __launch_bounds__(256) gene_insert_kernel(const int32_t *__restrict__ mgene,
                                                          const int32_t *__restrict__ msample,
                                                          const uint8_t *__restrict__ mclass, int64_t n_mut,
                                                          unsigned long long *keys, uint32_t *cnt, int64_t capacity,
                                                          int32_t *status)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_mut; i += stride) {
        const uint32_t cls = mclass[i];
        const int32_t g = mgene[i];
        if (cls > 4u || g < 0) continue;
        const unsigned long long key =
            (((unsigned long long)(uint32_t)g << 32) | (unsigned long long)(uint32_t)msample[i]) + 1ull;
        const int64_t slot = table_slot(keys, capacity, key);
        if (slot < 0) {
            *status = 1;
            continue;
        }
        atomicAdd(cnt + slot * 5 + cls, 1u);
    }
}

__global__ void __launch_bounds__(256) gene_finalize_kernel(const unsigned long long *__restrict__ keys,
                                                            const uint32_t *__restrict__ cnt, int64_t capacity,
                                                            int64_t cap_per, int64_t n_gene, int64_t *obs,
                                                            int64_t *nsamp)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < capacity; s += stride) {
        const unsigned long long key = keys[s];
        if (key == 0ull) continue;
        const int64_t g = (int64_t)((key - 1ull) >> 32);
        if (g >= n_gene) continue;                 // a named gene that is not in the model table
        int64_t c[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) c[j] = cnt[s * 5 + j];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int64_t v = c[j] > cap_per ? cap_per : c[j];
            if (v) atomicAdd(reinterpret_cast<unsigned long long *>(obs + 5 * g + j), (unsigned long long)v);
        }
        const bool flags[7] = {c[0] > 0, c[1] > 0, c[2] > 0, c[3] > 0, (c[2] + c[3]) > 0, (c[1] + c[2] + c[3]) > 0,
                               c[4] > 0};
#pragma unroll
        for (int j = 0; j < 7; ++j)
            if (flags[j]) atomicAdd(reinterpret_cast<unsigned long long *>(nsamp + 7 * g + j), 1ull);
    }
}

// L[elt, sub] += 1 per site (preprocess_sites, sequence_tools.py:698-703)
__global__ void __launch_bounds__(256) site_counts_kernel(const int32_t *__restrict__ elt, const int32_t *__restrict__ sub,
                                                          int64_t n, int64_t n_elt, int n_sub,
                                                          unsigned long long *__restrict__ L)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int32_t e = elt[i], j = sub[i];
        if (e >= 0 && e < n_elt && j >= 0 && j < n_sub) atomicAdd(L + (int64_t)e * n_sub + j, 1ull);
    }
}

inline unsigned grid_for(int64_t n)
{
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)dig::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}

// Overlap join (`bedtools intersect -wa -wb` itself, one output row per overlapping (mutation, block) pair) behind
// mutation_tools.restrict_mutations_by_bed(_efficient) (:8-43), mutations_by_element (:363-381) and
// tabulate_nonc_mutations_split (:120-153).  Two passes over the same stabbing walk: count per mutation, then (after
// the caller's exclusive scan) fill.  Within a mutation the pairs come out by ascending sorted-block index.
__global__ void __launch_bounds__(256) overlap_count_kernel(
    const int64_t *__restrict__ bks, const int64_t *__restrict__ bke, const int64_t *__restrict__ pmax, int64_t n_blk,
    const int64_t *__restrict__ mks, const int64_t *__restrict__ mke, int64_t n_mut, int64_t *__restrict__ n_pairs)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_mut; i += stride) {
        const int64_t ks = mks[i], ke = mke[i];
        int64_t c = 0;
        if (ke > ks) {
            const int64_t hi = upper_bound_ge(bks, n_blk, ke);
            for (int64_t idx = hi - 1; idx >= 0; --idx) {
                if (__ldg(pmax + idx) <= ks) break;
                c += __ldg(bke + idx) > ks;
            }
        }
        n_pairs[i] = c;
    }
}

__global__ void __launch_bounds__(256) overlap_fill_kernel(
    const int64_t *__restrict__ bks, const int64_t *__restrict__ bke, const int64_t *__restrict__ pmax, int64_t n_blk,
    const int64_t *__restrict__ mks, const int64_t *__restrict__ mke, int64_t n_mut,
    const int64_t *__restrict__ pair_off, int64_t *__restrict__ pair_mut, int64_t *__restrict__ pair_blk)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_mut; i += stride) {
        const int64_t ks = mks[i], ke = mke[i];
        if (ke <= ks) continue;
        int64_t w = pair_off[i + 1];                         // filled back to front: ascending block order
        const int64_t hi = upper_bound_ge(bks, n_blk, ke);
        for (int64_t idx = hi - 1; idx >= 0; --idx) {
            if (__ldg(pmax + idx) <= ks) break;
            if (__ldg(bke + idx) <= ks) continue;
            --w;
            pair_mut[w] = i;
            pair_blk[w] = idx;
        }
    }
}

inline bool is_pow2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace

extern "C" {

int64_t dig_tabulate_capacity(int64_t n_pairs)
{
    if (n_pairs < 0) n_pairs = 0;
    int64_t cap = 1024;
    while (cap < 2 * n_pairs + 16) cap <<= 1;
    return cap;
}
int64_t dig_tabulate_elements_workspace_bytes(int64_t n_pairs) { return dig_tabulate_capacity(n_pairs) * 16; }
int64_t dig_tabulate_genes_workspace_bytes(int64_t n_mut) { return dig_tabulate_capacity(n_mut) * 28; }

int dig_count_hits(const int64_t *blk_kstart_d, const int64_t *blk_kend_d, const int64_t *blk_pmax_d,
                   const int32_t *blk_elt_d, int64_t n_blk, const int64_t *mut_kstart_d,
                   const int64_t *mut_kend_d, int64_t n_mut, unsigned long long *n_hits_d, void *stream)
{
    DIG_CHECK_ARG(n_blk >= 0 && n_mut >= 0, "negative size");
    DIG_CHECK_ARG(n_hits_d != nullptr, "null pointer");
    if (n_blk == 0 || n_mut == 0) return DIG_OK;
    DIG_CHECK_ARG(blk_kstart_d && blk_kend_d && blk_pmax_d && blk_elt_d && mut_kstart_d && mut_kend_d, "null pointer");
    count_hits_kernel<<<grid_for(n_mut), 256, 0, (cudaStream_t)stream>>>(blk_kstart_d, blk_kend_d, blk_pmax_d,
                                                                         blk_elt_d, n_blk, mut_kstart_d, mut_kend_d,
                                                                         n_mut, n_hits_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

int dig_tabulate_elements(const int64_t *blk_kstart_d, const int64_t *blk_kend_d, const int64_t *blk_pmax_d,
                          const int32_t *blk_elt_d, int64_t n_blk, const int64_t *mut_kstart_d,
                          const int64_t *mut_kend_d, const int32_t *mut_sample_d, const uint8_t *mut_isindel_d,
                          int64_t n_mut, unsigned long long *tab_key_d, uint32_t *tab_snv_d, uint32_t *tab_indel_d,
                          int64_t capacity, int64_t n_sample, unsigned long long *sample_tot_d,
                          int64_t max_muts_per_sample, int64_t max_per_elt_per_sample, int64_t n_elt,
                          int64_t *obs_d, int32_t *status_d, int sample_rows_mode, void *stream)
{
    DIG_CHECK_ARG(n_blk >= 0 && n_mut >= 0 && n_elt >= 0 && n_sample >= 0, "negative size");
    DIG_CHECK_ARG(is_pow2(capacity), "capacity must be a power of two");
    DIG_CHECK_ARG(tab_key_d && tab_snv_d && tab_indel_d && status_d, "null pointer");
    DIG_CHECK_ARG((n_elt == 0 || obs_d) && (n_sample == 0 || sample_tot_d), "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    DIG_CUDA(cudaMemsetAsync(tab_key_d, 0, (size_t)capacity * sizeof(unsigned long long), st));
    DIG_CUDA(cudaMemsetAsync(tab_snv_d, 0, (size_t)capacity * sizeof(uint32_t), st));
    DIG_CUDA(cudaMemsetAsync(tab_indel_d, 0, (size_t)capacity * sizeof(uint32_t), st));
    DIG_CUDA(cudaMemsetAsync(status_d, 0, sizeof(int32_t), st));
    if (n_sample) DIG_CUDA(cudaMemsetAsync(sample_tot_d, 0, (size_t)n_sample * sizeof(unsigned long long), st));
    if (n_elt) DIG_CUDA(cudaMemsetAsync(obs_d, 0, (size_t)n_elt * 3 * sizeof(int64_t), st));
    if (n_blk == 0 || n_mut == 0) return DIG_OK;
    DIG_CHECK_ARG(blk_kstart_d && blk_kend_d && blk_pmax_d && blk_elt_d && mut_kstart_d && mut_kend_d &&
                      mut_sample_d && mut_isindel_d,
                  "null pointer");
    elt_insert_kernel<<<grid_for(n_mut), 256, 0, st>>>(blk_kstart_d, blk_kend_d, blk_pmax_d, blk_elt_d, n_blk,
                                                       mut_kstart_d, mut_kend_d, mut_sample_d, mut_isindel_d, n_mut,
                                                       tab_key_d, tab_snv_d, tab_indel_d, capacity, status_d);
    DIG_CHECK_LAUNCH();
    elt_sample_totals_kernel<<<grid_for(capacity), 256, 0, st>>>(tab_key_d, tab_snv_d, tab_indel_d, capacity,
                                                                 sample_rows_mode, sample_tot_d);
    DIG_CHECK_LAUNCH();
    elt_finalize_kernel<<<grid_for(capacity), 256, 0, st>>>(tab_key_d, tab_snv_d, tab_indel_d, capacity, sample_tot_d,
                                                            max_muts_per_sample, max_per_elt_per_sample, n_elt, obs_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

int dig_tabulate_genes(const int32_t *mut_gene_d, const int32_t *mut_sample_d, const uint8_t *mut_class_d,
                       int64_t n_mut, unsigned long long *tab_key_d, uint32_t *tab_cnt_d, int64_t capacity,
                       int64_t max_per_gene_per_sample, int64_t n_gene, int64_t *obs_d, int64_t *nsamp_d,
                       int32_t *status_d, void *stream)
{
    DIG_CHECK_ARG(n_mut >= 0 && n_gene >= 0, "negative size");
    DIG_CHECK_ARG(is_pow2(capacity), "capacity must be a power of two");
    DIG_CHECK_ARG(tab_key_d && tab_cnt_d && status_d && (n_gene == 0 || (obs_d && nsamp_d)), "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    DIG_CUDA(cudaMemsetAsync(tab_key_d, 0, (size_t)capacity * sizeof(unsigned long long), st));
    DIG_CUDA(cudaMemsetAsync(tab_cnt_d, 0, (size_t)capacity * 5 * sizeof(uint32_t), st));
    DIG_CUDA(cudaMemsetAsync(status_d, 0, sizeof(int32_t), st));
    if (n_gene) {
        DIG_CUDA(cudaMemsetAsync(obs_d, 0, (size_t)n_gene * 5 * sizeof(int64_t), st));
        DIG_CUDA(cudaMemsetAsync(nsamp_d, 0, (size_t)n_gene * 7 * sizeof(int64_t), st));
    }
    if (n_mut == 0) return DIG_OK;
    DIG_CHECK_ARG(mut_gene_d && mut_sample_d && mut_class_d, "null pointer");
    gene_insert_kernel<<<grid_for(n_mut), 256, 0, st>>>(mut_gene_d, mut_sample_d, mut_class_d, n_mut, tab_key_d,
                                                        tab_cnt_d, capacity, status_d);
    DIG_CHECK_LAUNCH();
    gene_finalize_kernel<<<grid_for(capacity), 256, 0, st>>>(tab_key_d, tab_cnt_d, capacity,
                                                             max_per_gene_per_sample, n_gene, obs_d, nsamp_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

int dig_site_counts(const int32_t *site_elt_d, const int32_t *site_sub_d, int64_t n_site, int64_t n_elt, int n_sub,
                    unsigned long long *L_d, void *stream)
{
    DIG_CHECK_ARG(n_site >= 0 && n_elt >= 0 && n_sub > 0, "bad sizes");
    DIG_CHECK_ARG(n_elt == 0 || L_d, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_elt) DIG_CUDA(cudaMemsetAsync(L_d, 0, (size_t)n_elt * n_sub * sizeof(unsigned long long), st));
    if (n_site == 0 || n_elt == 0) return DIG_OK;
    DIG_CHECK_ARG(site_elt_d && site_sub_d, "null pointer");
    site_counts_kernel<<<grid_for(n_site), 256, 0, st>>>(site_elt_d, site_sub_d, n_site, n_elt, n_sub, L_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

int dig_overlap_count(const int64_t *blk_kstart_d, const int64_t *blk_kend_d, const int64_t *blk_pmax_d, int64_t n_blk,
                      const int64_t *mut_kstart_d, const int64_t *mut_kend_d, int64_t n_mut, int64_t *n_pairs_d,
                      void *stream)
{
    DIG_CHECK_ARG(n_blk >= 0 && n_mut >= 0, "negative size");
    if (n_mut == 0) return DIG_OK;
    DIG_CHECK_ARG(mut_kstart_d && mut_kend_d && n_pairs_d, "null pointer");
    DIG_CHECK_ARG(n_blk == 0 || (blk_kstart_d && blk_kend_d && blk_pmax_d), "null pointer");
    overlap_count_kernel<<<grid_for(n_mut), 256, 0, (cudaStream_t)stream>>>(blk_kstart_d, blk_kend_d, blk_pmax_d, n_blk,
                                                                            mut_kstart_d, mut_kend_d, n_mut, n_pairs_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

int dig_overlap_fill(const int64_t *blk_kstart_d, const int64_t *blk_kend_d, const int64_t *blk_pmax_d, int64_t n_blk,
                     const int64_t *mut_kstart_d, const int64_t *mut_kend_d, int64_t n_mut, const int64_t *pair_off_d,
                     int64_t *pair_mut_d, int64_t *pair_blk_d, void *stream)
{
    DIG_CHECK_ARG(n_blk >= 0 && n_mut >= 0, "negative size");
    if (n_mut == 0 || n_blk == 0) return DIG_OK;
    DIG_CHECK_ARG(blk_kstart_d && blk_kend_d && blk_pmax_d && mut_kstart_d && mut_kend_d && pair_off_d, "null pointer");
    DIG_CHECK_ARG(pair_mut_d && pair_blk_d, "null pointer");
    overlap_fill_kernel<<<grid_for(n_mut), 256, 0, (cudaStream_t)stream>>>(blk_kstart_d, blk_kend_d, blk_pmax_d, n_blk,
                                                                           mut_kstart_d, mut_kend_d, n_mut, pair_off_d,
                                                                           pair_mut_d, pair_blk_d);
    DIG_CHECK_LAUNCH();
    return DIG_OK;
}

}
