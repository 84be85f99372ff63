/*
 * dig_oracle.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * A plain-C restatement of the integer / byte stages of DIGDriver's hot path,
 * written to follow the reference's Python line by line in *behaviour* (same
 * windows, same clipping quirks, same skip rules) so that the CUDA kernels in
 * digdriver_b200/csrc can be checked bit-exactly against it at sizes the pure
 * Python reference cannot reach.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * Parity pinning: the reference ships no tests or golden vectors (SURVEY.md
 * section 4), so this oracle is pinned against outputs of the UNMODIFIED
 * reference functions executed in the build container
 * (tests/golden/make_golden.py -> tests/golden/ npz files, checked by
 * tests/test_oracle_golden.py).
 *
 * Reference lines restated here (paths relative to /root/reference):
 *   orc_count_regions   DIGDriver/sequence_model/sequence_tools.py:21-29  (fetch_sequence)
 *                       :42-55 (seq_to_context) :65-78 (count_sequence_context)
 *                       :80-94 (count_contexts_by_regions)
 *                       :527-556 (nonc_elt_context_count: strand-aware variant)
 *   orc_mutation_contexts  sequence_tools.py:130-178 (mutation_contexts_by_chrom)
 *   orc_pack_genome     definition of the 2-bit + N-mask device layout (no
 *                       reference counterpart; the reference reads ASCII FASTA)
 *   orc_synth_genome    deterministic synthetic genome generator shared with the
 *                       CUDA generator (BASELINE.json configs are synthetic)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* helpers                                                             */
/* ------------------------------------------------------------------ */

/* upper-cased base -> 0..3, anything else (N, IUPAC, pad) -> 4 */
static inline int base_code(uint8_t c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}

static inline int is_n_char(uint8_t c) { return c == 'N' || c == 'n'; }

/* ------------------------------------------------------------------ */
/* context counting                                                    */
/* ------------------------------------------------------------------ */

/*
 * Count (n_up+1+n_down)-mers centred on every base of each region, exactly as
 * count_contexts_by_regions does:
 *   - START == 0 is silently replaced by n_up            (sequence_tools.py:25-26)
 *   - the fetched string is [START-n_up, END+n_down) clipped at the chromosome
 *     end, so the last n_down centres of a chromosome are never counted (:28,:71)
 *   - a k-mer containing 'N' is skipped                  (:48-49, :73-74)
 *   - strand < 0: the fetched string is reverse-complemented first (:552-553)
 * Column order: lexicographic over ACGT with the 5' base most significant
 * (itertools.product order, sequence_tools.py:37-40).
 *
 * seq        concatenated ASCII chromosomes (mixed case allowed; upper-cased on read)
 * chrom_off  [n_chrom+1] start offset of each chromosome in seq; chromosome c has
 *            length chrom_len[c]
 * counts     [n_reg, K] int64, K = 4^(n_up+1+n_down); overwritten
 * n_other    (optional) number of k-mers skipped because of a non-ACGTN byte;
 *            the reference would raise KeyError on those (quirk a1-iv)
 * returns 0, or -1 for a region with 0 < START < n_up (pysam raises there)
 */
int orc_count_regions(const uint8_t *seq, const int64_t *chrom_off, const int64_t *chrom_len,
                      const int32_t *reg_chrom, const int64_t *reg_start, const int64_t *reg_end,
                      const int8_t *reg_strand, int64_t n_reg, int n_up, int n_down,
                      int64_t *counts, int64_t *n_other)
{
    const int klen = n_up + 1 + n_down;
    const int64_t K = (int64_t)1 << (2 * klen);
    int64_t other_total = 0;
    int err = 0;

#pragma omp parallel for schedule(dynamic, 64) reduction(+ : other_total)
    for (int64_t r = 0; r < n_reg; ++r) {
        int64_t *out = counts + r * K;
        memset(out, 0, (size_t)K * sizeof(int64_t));
        const int32_t c = reg_chrom[r];
        const int64_t L = chrom_len[c];
        const uint8_t *s = seq + chrom_off[c];
        int64_t start = reg_start[r];
        const int64_t end = reg_end[r];
        if (start == 0) start = n_up;
        if (start < n_up) {
#pragma omp atomic write
            err = -1;
            continue;
        }
        int64_t f0 = start - n_up;          /* fetched [f0, f1) */
        int64_t f1 = end + n_down;
        if (f1 > L) f1 = L;
        if (f0 > L) f0 = L;
        const int64_t n = f1 - f0;
        if (n < klen) continue;
        const int minus = reg_strand && reg_strand[r] < 0;
        /* walk the fetched string the way the reference walks it: i is the index
         * of the centre in the (possibly reverse-complemented) string */
        for (int64_t i = n_up; i < n - n_down; ++i) {
            int64_t key = 0;
            int bad = 0, other = 0;
            for (int t = -n_up; t <= n_down; ++t) {
                int code;
                uint8_t ch;
                if (!minus) {
                    ch = s[f0 + i + t];
                    code = base_code(ch);
                } else {
                    ch = s[f0 + (n - 1 - (i + t))];
                    code = base_code(ch);
                    if (code < 4) code = 3 - code;
                }
                if (code > 3) {
                    bad = 1;
                    if (!is_n_char(ch)) other = 1;
                }
                key = (key << 2) | (code & 3);
            }
            if (bad) {
                other_total += other;
                continue;
            }
            out[key] += 1;
        }
    }
    if (n_other) *n_other = other_total;
    return err;
}

/* ------------------------------------------------------------------ */
/* mutation contexts                                                   */
/* ------------------------------------------------------------------ */

/*
 * mutation_contexts_by_chrom (sequence_tools.py:130-178), one chromosome group
 * at a time in file order.  ref/alt are 0..3 for a single A/C/G/T character and
 * anything else (multi-base, N, lower case) is 255 -- such a REF never equals
 * the upper-cased genome character, so the row is dropped (:145-148).
 *
 * Quirk reproduced (:150-151): a row whose START equals the previous row's START
 * re-uses the previous row's CONTEXT, so once one row of a same-START run is
 * dropped for a REF mismatch every later row of that run is dropped too.
 *
 * ctx_out[i]  k-mer index (5' base most significant) or -1 when the reference
 *             drops the row (REF mismatch, N in context, inherited drop).  Rows
 *             whose context would run off either chromosome end are dropped
 *             (the reference yields '' at the left edge and a truncated string at
 *             the right edge; we do not reproduce truncated strings).
 * grp_start   rows [grp_start[g], grp_start[g+1]) share one chromosome
 */
int orc_mutation_contexts(const uint8_t *seq, const int64_t *chrom_off, const int64_t *chrom_len,
                          const int32_t *mut_chrom, const int64_t *mut_start,
                          const uint8_t *mut_ref, int64_t n_mut,
                          int n_up, int n_down, int32_t *ctx_out)
{
    int64_t prev_start = -1;
    int32_t prev_chrom = -1;
    int32_t prev_ctx = -1;
    for (int64_t i = 0; i < n_mut; ++i) {
        const int32_t c = mut_chrom[i];
        const int64_t L = chrom_len[c];
        const uint8_t *s = seq + chrom_off[c];
        const int64_t p = mut_start[i];
        int32_t ctx;
        if (c != prev_chrom) prev_start = -1;   /* a new chromosome group starts a new loop */
        if (p < 0 || p >= L || mut_ref[i] > 3 || base_code(s[p]) != mut_ref[i]) {
            ctx = -1;
        } else if (p == prev_start) {
            ctx = prev_ctx;
        } else if (p - n_up < 0 || p + n_down >= L) {
            ctx = -1;
        } else {
            int32_t key = 0;
            int bad = 0;
            for (int t = -n_up; t <= n_down; ++t) {
                int code = base_code(s[p + t]);
                if (code > 3) bad = 1;
                key = (key << 2) | (code & 3);
            }
            ctx = bad ? -1 : key;
        }
        ctx_out[i] = ctx;
        prev_start = p;
        prev_chrom = c;
        prev_ctx = ctx;
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* device layout definition: 2-bit bases + N bitmask, MSB first        */
/* ------------------------------------------------------------------ */

/*
 * packed2[g >> 4] holds base g in bits [30 - 2*(g&15), 31 - 2*(g&15)]  (A=0 C=1 G=2 T=3,
 * anything else stored as 0); nmask[g >> 5] bit (31 - (g&31)) is 1 when base g is
 * not A/C/G/T.  g is the global coordinate chrom_off[c] + position.  n is the number
 * of global positions; words beyond it are padded as N.
 */
void orc_pack_genome(const uint8_t *seq, int64_t n, uint32_t *packed2, uint32_t *nmask)
{
    const int64_t nw2 = (n + 15) >> 4;
    const int64_t nwn = (n + 31) >> 5;
#pragma omp parallel for schedule(static)
    for (int64_t w = 0; w < nw2; ++w) {
        uint32_t v = 0;
        for (int i = 0; i < 16; ++i) {
            int64_t g = (w << 4) + i;
            int code = g < n ? base_code(seq[g]) : 4;
            v |= (uint32_t)(code & 3) << (30 - 2 * i);
        }
        packed2[w] = v;
    }
#pragma omp parallel for schedule(static)
    for (int64_t w = 0; w < nwn; ++w) {
        uint32_t v = 0;
        for (int i = 0; i < 32; ++i) {
            int64_t g = (w << 5) + i;
            int code = g < n ? base_code(seq[g]) : 4;
            if (code > 3) v |= 1u << (31 - i);
        }
        nmask[w] = v;
    }
}

/* ------------------------------------------------------------------ */
/* synthetic genome (shared definition with csrc/synth.cu)             */
/* ------------------------------------------------------------------ */

static inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/*
 * Position g of the synthetic genome: i.i.d. ACGT (p = 1/4), 50 % lower case, and one
 * N run of 10-50 kb per 1 Mb super-block (about 3 % N), all pure functions of
 * (seed, g) so that any slice can be generated anywhere.  n_frac16 scales the N run
 * length in 1/16ths (16 = as described, 0 = no N at all).
 */
void orc_synth_genome(uint8_t *seq, int64_t g0, int64_t n, uint64_t seed, int n_frac16)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t g = (uint64_t)(g0 + i);
        const uint64_t h = mix64(seed ^ (g * 0xD1342543DE82EF95ull));
        uint8_t ch = "ACGT"[h & 3];
        if ((h >> 2) & 1) ch = (uint8_t)(ch | 0x20);
        const uint64_t sb = g >> 20;
        const uint64_t hs = mix64(seed ^ 0xA5A5A5A5ull ^ (sb * 0x9E3779B97F4A7C15ull));
        const uint64_t off = hs % (uint64_t)((1 << 20) - 51200);
        const uint64_t len = ((10240 + (hs >> 32) % 40960) * (uint64_t)n_frac16) >> 4;
        const uint64_t q = g & ((1u << 20) - 1);
        if (q >= off && q < off + len) ch = 'N';
        seq[i] = ch;
    }
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
