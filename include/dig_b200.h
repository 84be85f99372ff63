/*
 * dig_b200.h -- C ABI of libdigb200.so: the B200 (sm_100a) kernels behind DIGDriver's
 * genome-scan -> element-transfer -> burden-test hot path.
 *
 * The reference (maxwellsh/DIGDriver) is pure Python and has no FFI of its own; the
 * functions below are what a maintainer would bind (ctypes) in
 * place of the Python loops cited next to each entry point.  Paths are relative to the
 * reference root.
 *
 * Conventions
 *   - every pointer with a _d suffix is DEVICE memory owned by the caller; the library
 *     never allocates or frees caller-visible memory and keeps no global state except a
 *     thread-local error string;
 *   - all work is enqueued on the caller's `stream` (a cudaStream_t passed as void*) and
 *     the call returns without synchronising unless stated otherwise;
 *   - return value: 0 on success, a negative DIG_ERR_* code otherwise; dig_last_error()
 *     gives a message for the calling thread.  No exceptions, no exit().
 *   - k-mer / context index: base-4 number over A=0,C=1,G=2,T=3 with the 5' base most
 *     significant, i.e. the itertools.product column order of
 *     DIGDriver/sequence_model/sequence_tools.py:31-40.
 *   - genome coordinates are "global": g = chrom_off[c] + position, chromosomes are
 *     concatenated in one buffer (the host aligns chrom_off to 128 bases and pads with N).
 */
#ifndef DIG_B200_H
#define DIG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIG_OK 0
#define DIG_ERR_ARG (-1)      /* bad argument (null pointer, negative size, unsupported n_up/n_down) */
#define DIG_ERR_CUDA (-2)     /* a CUDA runtime call or launch failed; see dig_last_error() */
#define DIG_ERR_RANGE (-3)    /* coordinates outside the genome */
#define DIG_ERR_UNSUPPORTED (-4)

int dig_version(void);
const char *dig_last_error(void);
/* number of SMs of the current device (grid sizing is done inside the library) */
int dig_device_sm_count(void);

/* ---------------------------------------------------------------------------------
 * K1  ASCII genome -> 2-bit bases + N bitmask.
 * Replaces: pysam fetch + str.upper() of sequence_tools.py:28 / :136 as the way the
 * genome reaches the counting loops.
 *   ascii_d    [n] bytes, mixed case, anything not ACGT (either case) is "N-like"
 *   packed2_d  [dig_packed_words(n)] uint32, base g in bits [30-2*(g&15), 31-2*(g&15)] of word g>>4
 *   nmask_d    [dig_nmask_words(n)] uint32, bit 31-(g&31) of word g>>5 set when base g is not ACGT
 *   n_other_d  (nullable) += number of bytes that are neither ACGT nor N (the reference raises
 *              KeyError on those at sequence_tools.py:76)
 * Positions >= n inside the last words are flagged N.
 */
int64_t dig_packed_words(int64_t n_bases);
int64_t dig_nmask_words(int64_t n_bases);
int dig_pack_genome(const uint8_t *ascii_d, int64_t n_bases, uint32_t *packed2_d, uint32_t *nmask_d,
                    unsigned long long *n_other_d, void *stream);

/* ---------------------------------------------------------------------------------
 * K2 / K4  per-region context histogram.
 * Replaces: count_sequence_context / count_contexts_by_regions / count_contexts_in_bed
 * (sequence_tools.py:65-128) and, with reg_strand_d, nonc_elt_context_count (:527-556).
 * Semantics reproduced bit-exactly: START==0 -> n_up (:25-26); clipping at the chromosome
 * end (:28,:71); k-mers containing a non-ACGT base are skipped (:48-49); strand < 0 counts
 * the reverse-complemented string (:552-553).
 *   reg_chrom_d  [n_reg] index into chrom_off_d / chrom_len_d
 *   reg_start_d, reg_end_d  [n_reg] 0-based half-open chromosome coordinates
 *   reg_strand_d (nullable) [n_reg] int8, < 0 = minus strand
 *   counts_d     [n_reg, K] int32, K = 4^(n_up+1+n_down); fully overwritten
 *   totals_d     (nullable) [K] uint64, column sums are ADDED to it (zero it first): the
 *                genome-wide totals of DigPreprocess.py:59 fused into the scan
 * Supported: n_up + n_down <= 5 (K <= 4096).  A region with 0 < START < n_up is counted
 * from n_up (the reference raises inside pysam); hosts should reject it beforehand.
 */
int dig_count_contexts(const uint32_t *packed2_d, const uint32_t *nmask_d, int64_t n_bases,
                       const int64_t *chrom_off_d, const int64_t *chrom_len_d,
                       const int32_t *reg_chrom_d, const int64_t *reg_start_d, const int64_t *reg_end_d,
                       const int8_t *reg_strand_d, int64_t n_reg, int n_up, int n_down,
                       int32_t *counts_d, unsigned long long *totals_d, void *stream);

/* ---------------------------------------------------------------------------------
 * Synthetic genome generator (BASELINE.json configs are synthetic): position g is a pure
 * function of (seed, g); identical to orc_synth_genome in oracle/dig_oracle.c.
 */
int dig_synth_genome(uint8_t *ascii_d, int64_t g0, int64_t n, uint64_t seed, int n_frac16, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DIG_B200_H */
