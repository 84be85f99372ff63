"""TEST INFRASTRUCTURE -- CPU oracle for the DIGDriver hot path.  NOT product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product package
``digdriver_b200`` never does (tests/test_boundary.py greps for it).

What is here:
  * ctypes wrappers over ``oracle/dig_oracle.c`` (integer/byte stages, OpenMP);
  * numpy / pandas / scipy restatements of the stages the reference implements
    with h5py- or bedtools-bound loops, each citing the reference lines it
    follows (paths relative to /root/reference);
  * the p-value arithmetic itself is the reference's own expression evaluated
    with SciPy (third-party, not under /root/reference: scipy.special.betainc,
    scipy.stats.nbinom.pmf, scipy.stats.chi2.sf; reference pins scipy 1.5.3 in
    conda-recipe/meta.yaml:45, this image has 1.18.1 -- both implement the same
    mathematical functions to ~1e-14).

Parity pinning: every function below is checked against outputs of the
UNMODIFIED reference functions executed in the build container
(tests/golden/make_golden.py -> tests/golden/*.npz; tests/test_oracle_golden.py).
"""
import ctypes
import itertools
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libdig_oracle.so")
    src = os.path.join(_HERE, "dig_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libdig_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libdig_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
        _LIB.orc_count_regions.restype = ctypes.c_int
        _LIB.orc_mutation_contexts.restype = ctypes.c_int
        _LIB.orc_num_threads.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


# ----------------------------------------------------------------------------
# integer stages (C)
# ----------------------------------------------------------------------------

def count_regions(seq, chrom_off, chrom_len, reg_chrom, reg_start, reg_end,
                  n_up=1, n_down=1, strand=None, threads=None):
    """count_contexts_by_regions / nonc_elt_context_count (sequence_tools.py:80-94, :527-556).

    Returns (counts int64 [n, K], n_other)."""
    L = lib()
    if threads:
        L.orc_set_threads(int(threads))
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    chrom_off = np.ascontiguousarray(chrom_off, dtype=np.int64)
    chrom_len = np.ascontiguousarray(chrom_len, dtype=np.int64)
    reg_chrom = np.ascontiguousarray(reg_chrom, dtype=np.int32)
    reg_start = np.ascontiguousarray(reg_start, dtype=np.int64)
    reg_end = np.ascontiguousarray(reg_end, dtype=np.int64)
    st = None if strand is None else np.ascontiguousarray(strand, dtype=np.int8)
    n = len(reg_chrom)
    K = 4 ** (n_up + n_down + 1)
    counts = np.empty((n, K), dtype=np.int64)
    n_other = ctypes.c_int64(0)
    rc = L.orc_count_regions(_p(seq), _p(chrom_off), _p(chrom_len), _p(reg_chrom), _p(reg_start),
                             _p(reg_end), _p(st), ctypes.c_int64(n), ctypes.c_int(n_up),
                             ctypes.c_int(n_down), _p(counts), ctypes.byref(n_other))
    if rc != 0:
        raise ValueError("start out of range (0 < START < n_up)")
    return counts, n_other.value


def mutation_contexts(seq, chrom_off, chrom_len, mut_chrom, mut_start, mut_ref, n_up=1, n_down=1):
    """mutation_contexts_by_chrom (sequence_tools.py:130-178); rows must be grouped by
    chromosome (stable), as pandas groupby delivers them.  Returns ctx int32 (-1 = dropped)."""
    L = lib()
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    chrom_off = np.ascontiguousarray(chrom_off, dtype=np.int64)
    chrom_len = np.ascontiguousarray(chrom_len, dtype=np.int64)
    mut_chrom = np.ascontiguousarray(mut_chrom, dtype=np.int32)
    mut_start = np.ascontiguousarray(mut_start, dtype=np.int64)
    mut_ref = np.ascontiguousarray(mut_ref, dtype=np.uint8)
    out = np.empty(len(mut_chrom), dtype=np.int32)
    L.orc_mutation_contexts(_p(seq), _p(chrom_off), _p(chrom_len), _p(mut_chrom), _p(mut_start),
                            _p(mut_ref), ctypes.c_int64(len(mut_chrom)), ctypes.c_int(n_up),
                            ctypes.c_int(n_down), _p(out))
    return out


def pack_genome(seq):
    """Definition of the device layout (2-bit MSB-first + N bitmask)."""
    L = lib()
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    n = len(seq)
    packed = np.empty((n + 15) // 16, dtype=np.uint32)
    nmask = np.empty((n + 31) // 32, dtype=np.uint32)
    L.orc_pack_genome(_p(seq), ctypes.c_int64(n), _p(packed), _p(nmask))
    return packed, nmask


def synth_genome(g0, n, seed, n_frac16=16):
    L = lib()
    seq = np.empty(n, dtype=np.uint8)
    L.orc_synth_genome(_p(seq), ctypes.c_int64(g0), ctypes.c_int64(n), ctypes.c_uint64(seed),
                       ctypes.c_int(n_frac16))
    return seq


def num_threads():
    return lib().orc_num_threads()


# ----------------------------------------------------------------------------
# pure-Python line-faithful port of the per-base loop (small cases + the
# reference-arm timing in bench.py: same language and cost model as the reference)
# ----------------------------------------------------------------------------

def py_count_sequence_context(seq, n_up=2, n_down=2):
    """count_sequence_context (sequence_tools.py:65-78) on an upper-cased str."""
    letters = "ACGT"
    table = {"".join(t): 0 for t in itertools.product(letters, repeat=n_up + 1 + n_down)}
    for i in range(n_up, len(seq) - n_down):
        kmer = seq[i - n_up:i + n_down + 1]
        if "N" in kmer:
            continue
        table[kmer] += 1
    return table


def py_count_regions(chrom_seqs, chroms, starts, ends, n_up=2, n_down=2):
    """count_contexts_by_regions (sequence_tools.py:80-94) with fetch_sequence (:21-29).

    chrom_seqs: {name: str}.  Returns int64 [n, K] in itertools.product column order."""
    rows = []
    for c, s, e in zip(chroms, starts, ends):
        if s == 0:
            s = n_up
        seq = chrom_seqs[c][s - n_up:e + n_down].upper()
        rows.append(list(py_count_sequence_context(seq, n_up, n_down).values()))
    return np.array(rows, dtype=np.int64)


def _py_chunk(args):
    return py_count_regions(*args)


def py_count_regions_pool(chrom_seqs, chroms, starts, ends, n_up, n_down, n_proc):
    """count_contexts_in_bed's multiprocessing.Pool chunking (sequence_tools.py:103-125)."""
    import multiprocessing as mp
    n = len(chroms)
    chunk = max(1, int(n / n_proc))
    bounds = list(range(0, n, chunk)) + [n]
    jobs = [(chrom_seqs, chroms[a:b], starts[a:b], ends[a:b], n_up, n_down)
            for a, b in zip(bounds[:-1], bounds[1:])]
    with mp.Pool(n_proc) as pool:
        parts = pool.map(_py_chunk, jobs)
    return np.concatenate(parts, axis=0)


# ----------------------------------------------------------------------------
# index tables
# ----------------------------------------------------------------------------

BASES = "ACGT"
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def revcomp(s):
    return "".join(_COMP[c] for c in reversed(s))


def context_names(n_up=1, n_down=1):
    """mk_context_sequences key order (sequence_tools.py:31-40)."""
    return ["".join(t) for t in itertools.product(BASES, repeat=n_up + 1 + n_down)]


def mutation_context_table():
    """mk_mutation_context(1,1) row order (sequence_tools.py:232-267): list of (MUT_TYPE, CONTEXT)."""
    rows = []
    for ref, alts in (("A", "TCG"), ("C", "AGT"), ("G", "TCA"), ("T", "AGC")):
        ctxs = [a + ref + b for a in BASES for b in BASES]
        for alt in alts:
            for ctx in ctxs:
                rows.append(("%s>%s" % (ref, alt), ctx))
    return rows


def substitution_names():
    """mk_trans_idx(1,1) (sequence_tools.py:282-289): sorted 'CTX>CTX2' strings, 192 of them."""
    return sorted(c + ">" + c[0] + m[2] + c[2] for m, c in mutation_context_table())


def model192_to_dpr(freq_192):
    """d_pr: FREQ re-ordered from mk_mutation_context row order to sorted substitution order
    (genic_driver_tools.py:321-325)."""
    names = [c + ">" + c[0] + m[2] + c[2] for m, c in mutation_context_table()]
    order = np.argsort(np.array(names), kind="stable")
    return np.asarray(freq_192, dtype=np.float64)[order]


def revcomp_permutation_192():
    """perm such that new[i] = old[perm[i]] reproduces the reference's minus-strand
    re-ordering of region_counts (sequence_tools.py:612-614, :633-634)."""
    names = substitution_names()
    pos = {n: i for i, n in enumerate(names)}
    rc = [revcomp(n.split(">")[0]) + ">" + revcomp(n.split(">")[1]) for n in names]
    # sorted(enumerate(x), key=rc_name_of_index): the i-th output is the old entry whose
    # rc-name is the i-th smallest, i.e. whose rc-name == names[i]
    return np.array([rc.index(names[i]) for i in range(192)], dtype=np.int64), pos


# ----------------------------------------------------------------------------
# element pretrain (stage 2)
# ----------------------------------------------------------------------------

def ideal_overlaps(block_starts, block_ends, window):
    """get_ideal_overlaps (genic_driver_tools.py:275-283) -> sorted unique window starts."""
    out = set()
    for s, e in zip(block_starts, block_ends):
        low = math.floor(s / window) * window
        high = math.ceil(e / window) * window
        borders = np.arange(low, high + window, window)
        for i in range(len(borders) - 1):
            out.add(int(borders[i]))
    return sorted(out)


def element_transfer(elt_chrom, elt_strand, blk_ptr, blk_start, blk_end, L192,
                     window, win_index, win_counts64, y_pred, std, y_true, flag, d_pr):
    """The loop body of DIG_onthefly (onthefly_tools.py:109-165) == preprocess_nonc
    (sequence_tools.py:619-641) + nonc_model (genic_driver_tools.py:347-390).

    win_index: dict {(chrom, window_start): row}.  L192: float [E, 192] (already strand-aware).
    Returns dict of per-element arrays."""
    E = len(elt_chrom)
    perm, _ = revcomp_permutation_192()
    out = {k: np.zeros(E) for k in ("MU", "SIGMA", "R_OBS", "P_SUM", "P_INDEL")}
    out["FLAG"] = np.zeros(E, dtype=bool)
    out["R_SIZE"] = np.zeros(E, dtype=np.int64)
    out["ELT_SIZE"] = np.zeros(E, dtype=np.int64)
    out["N_WIN"] = np.zeros(E, dtype=np.int64)
    d_pr = np.asarray(d_pr, dtype=np.float64)
    for i in range(E):
        a, b = blk_ptr[i], blk_ptr[i + 1]
        wins = ideal_overlaps(blk_start[a:b], blk_end[a:b], window)
        rows = [win_index[(int(elt_chrom[i]), w)] for w in wins]
        out["N_WIN"][i] = len(rows)
        region_counts = np.zeros(192)
        for r in rows:
            region_counts = region_counts + np.repeat(win_counts64[r], 3)
        if elt_strand[i] < 0:
            region_counts = region_counts[perm]
        L = np.asarray(L192[i], dtype=np.float64)
        prob_sum = region_counts * d_pr
        with np.errstate(divide="ignore", invalid="ignore"):
            t_pi = d_pr / prob_sum.sum()
            out["P_SUM"][i] = (t_pi * L).sum()
        mu = 0.0
        var = 0.0
        robs = 0.0
        fl = False
        for r in rows:                      # get_region_params_direct (genic_driver_tools.py:258-272)
            mu += y_pred[r]
            var += std[r] ** 2
            robs += y_true[r]
            fl = fl or bool(flag[r])
        out["MU"][i] = mu
        out["SIGMA"][i] = np.sqrt(var)
        out["R_OBS"][i] = robs
        out["FLAG"][i] = fl
        out["R_SIZE"][i] = int(region_counts.sum() / 3)
        out["ELT_SIZE"][i] = int(np.sum(L) / 3)
        with np.errstate(divide="ignore", invalid="ignore"):
            out["P_INDEL"][i] = np.float64(out["ELT_SIZE"][i]) / np.float64(out["R_SIZE"][i])
    return out


def gene_transfer(gene_chrom, gene_strand, blk_ptr, blk_start, blk_end, L192x4,
                  window, win_index, win_counts64, y_pred, std, y_true, flag, d_pr):
    """genic_model (genic_driver_tools.py:86-168) with the precounted region contexts of
    si_count_pretrain / si_by_regions (sequence_tools.py:375-425) folded in.

    L192x4: [E, 192, 4] columns (silent, mis, nons, splice).  Gene intervals are the
    inclusive CDS intervals of f_genic (GENE_LENGTH = sum(end - start + 1), :158)."""
    E = len(gene_chrom)
    base = element_transfer(gene_chrom, gene_strand, blk_ptr, blk_start, blk_end,
                            np.zeros((E, 192)), window, win_index, win_counts64,
                            y_pred, std, y_true, flag, d_pr)
    perm, _ = revcomp_permutation_192()
    P = np.zeros((E, 4))
    glen = np.zeros(E, dtype=np.int64)
    d_pr = np.asarray(d_pr, dtype=np.float64)
    for i in range(E):
        a, b = blk_ptr[i], blk_ptr[i + 1]
        wins = ideal_overlaps(blk_start[a:b], blk_end[a:b], window)
        rows = [win_index[(int(gene_chrom[i]), w)] for w in wins]
        rc = np.zeros(192)
        for r in rows:
            rc = rc + np.repeat(win_counts64[r], 3)
        if gene_strand[i] < 0:
            rc = rc[perm]
        with np.errstate(divide="ignore", invalid="ignore"):
            t_pi = d_pr / (rc * d_pr).sum()
            P[i] = (t_pi[:, None] * np.asarray(L192x4[i], dtype=np.float64)).sum(axis=0)
        glen[i] = int(np.sum(np.asarray(blk_end[a:b]) - np.asarray(blk_start[a:b]) + 1))
    out = {k: base[k] for k in ("MU", "SIGMA", "R_OBS", "FLAG", "R_SIZE", "N_WIN")}
    out["P_SILENT"], out["P_MIS"], out["P_NONS"], out["P_SPLICE"] = P.T
    out["P_TRUNC"] = out["P_NONS"] + out["P_SPLICE"]
    out["GENE_LENGTH"] = glen
    with np.errstate(divide="ignore", invalid="ignore"):
        out["P_INDEL"] = glen / out["R_SIZE"].astype(np.float64)
    return out


# ----------------------------------------------------------------------------
# the test (stage 3)
# ----------------------------------------------------------------------------

def normal_params_to_gamma(mu, sigma):
    """nb_model.py:237-241."""
    mu = np.asarray(mu, dtype=np.float64)
    sigma = np.asarray(sigma, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return mu ** 2 / sigma ** 2, sigma ** 2 / mu


def nb_pvalue_greater_midp(k, alpha, p):
    """nb_model.py:271-278 -- the reference's expression, evaluated with SciPy."""
    import scipy.special
    import scipy.stats
    return 0.5 * scipy.stats.nbinom.pmf(k, alpha, p) + scipy.special.betainc(np.asarray(k) + 1, alpha, 1 - np.asarray(p))


def burden_test(k, alpha, theta, pi):
    """element_expected_muts_nb + element_pvalue_burden_nb (transfer_tools.py:343-344, :473-482):
    EXP = ALPHA*THETA*Pi ; p = nb_pvalue_greater_midp(k, ALPHA, 1/(THETA*Pi+1))."""
    k = np.asarray(k, dtype=np.float64)
    alpha = np.asarray(alpha, dtype=np.float64)
    theta = np.asarray(theta, dtype=np.float64)
    pi = np.asarray(pi, dtype=np.float64)
    with np.errstate(all="ignore"):
        exp = alpha * theta * pi
        pval = nb_pvalue_greater_midp(k, alpha, 1 / (theta * pi + 1))
    return exp, pval


def fisher2(p1, p2):
    """transfer_tools.py:860-861, :1086-1087."""
    import scipy.stats
    with np.errstate(all="ignore"):
        x2 = -2 * (np.log(p1) + np.log(p2))
        return scipy.stats.chi2.sf(x2, df=4)


# ----------------------------------------------------------------------------
# observed counts (stage 0/3): pandas restatement of the bedtools-bound code
# ----------------------------------------------------------------------------

def tabulate_mutations_in_element(df_mut, df_blocks, max_muts_per_sample=1e9,
                                  max_muts_per_elt_per_sample=3e9, drop_duplicates=True):
    """tabulate_muts_per_sample_per_element + tabulate_mutations_in_element
    (mutation_tools.py:191-230, :155-189).

    df_mut: raw mutation rows in file order with columns CHROM START END REF ALT SAMPLE ANNOT.
    df_blocks: bed6 rows CHROM START END ELT (bed12tobed6 output).
    bedtools ``intersect -wa -wb`` is restated as: rows pair up when CHROM is equal and
    [START,END) and [bs,be) share at least one base."""
    import pandas as pd
    hits = []
    for chrom, dm in df_mut.groupby("CHROM", sort=False):
        db = df_blocks[df_blocks.CHROM == chrom]
        if len(db) == 0:
            continue
        bs = db.START.values
        be = db.END.values
        elt = db.ELT.values
        for row in dm.itertuples(index=True):
            m = (bs < row.END) & (be > row.START)
            for e in elt[m]:
                hits.append((row.Index, e))
    empty = pd.DataFrame({"OBS_SAMPLES": [], "OBS_SNV": [], "OBS_INDEL": [], "ELT": []}).set_index("ELT")
    if not hits:
        return empty, []
    h = pd.DataFrame(hits, columns=["ROW", "ELT"])
    df = df_mut.loc[h.ROW.values].reset_index(drop=True)
    df["ELT"] = h.ELT.values
    if drop_duplicates:
        df = df.drop_duplicates(["CHROM", "START", "END", "REF", "ALT", "SAMPLE", "ELT"])
    is_indel = df.ANNOT == "INDEL"
    snv = df[~is_indel].groupby(["ELT", "SAMPLE"]).size().reset_index(name="OBS_SNV")
    ind = df[is_indel].groupby(["ELT", "SAMPLE"]).size().reset_index(name="OBS_INDEL")
    cnt = snv.merge(ind, how="outer")
    cnt["OBS_SNV"] = cnt.OBS_SNV.fillna(0)
    cnt["OBS_INDEL"] = cnt.OBS_INDEL.fillna(0)
    cnt["OBS_MUT"] = cnt.OBS_SNV + cnt.OBS_INDEL
    per_sample = cnt.groupby("SAMPLE").OBS_MUT.sum()
    blacklist = list(per_sample[per_sample > max_muts_per_sample].index)
    cnt = cnt[~cnt.SAMPLE.isin(blacklist)].copy()
    cnt.loc[cnt.OBS_SNV > max_muts_per_elt_per_sample, "OBS_SNV"] = max_muts_per_elt_per_sample
    cnt.loc[cnt.OBS_INDEL > max_muts_per_elt_per_sample, "OBS_INDEL"] = max_muts_per_elt_per_sample
    if len(cnt) == 0:
        return empty, blacklist
    summ = cnt.groupby("ELT").agg(OBS_SAMPLES=("SAMPLE", "size"), OBS_SNV=("OBS_SNV", "sum"),
                                  OBS_INDEL=("OBS_INDEL", "sum"))
    return summ[["OBS_SAMPLES", "OBS_SNV", "OBS_INDEL"]], blacklist


def gene_observed_counts(df_mut_cds, max_muts_per_gene_per_sample=3e9):
    """mutations_per_gene (mutation_tools.py:329-361) + the distinct-sample counts of
    transfer_gene_model (transfer_tools.py:235-265).  Returns DataFrame indexed by GENE."""
    import pandas as pd
    g = df_mut_cds.groupby(["GENE", "SAMPLE", "ANNOT"]).size().reset_index(name="COUNT")
    capped = g.copy()
    capped.loc[capped.COUNT > max_muts_per_gene_per_sample, "COUNT"] = max_muts_per_gene_per_sample
    tab = capped.pivot_table(index="GENE", columns="ANNOT", values="COUNT", aggfunc="sum").fillna(0)
    names = {"Synonymous": "OBS_SYN", "Missense": "OBS_MIS", "Nonsense": "OBS_NONS",
             "Essential_Splice": "OBS_SPL", "INDEL": "OBS_INDEL"}
    out = pd.DataFrame(index=tab.index)
    for a, col in names.items():
        out[col] = tab[a].astype(np.int64) if a in tab.columns else 0
    classes = {"N_SAMP_SYN": ["Synonymous"], "N_SAMP_MIS": ["Missense"], "N_SAMP_NONS": ["Nonsense"],
               "N_SAMP_SPL": ["Essential_Splice"], "N_SAMP_TRUNC": ["Nonsense", "Essential_Splice"],
               "N_SAMP_NONSYN": ["Missense", "Nonsense", "Essential_Splice"], "N_SAMP_INDEL": ["INDEL"]}
    for col, annots in classes.items():
        sub = df_mut_cds[df_mut_cds.ANNOT.isin(annots)]
        ns = sub.groupby(["GENE", "SAMPLE"]).size().reset_index(name="CNT").GENE.value_counts()
        out[col] = ns.reindex(out.index).fillna(0).astype(np.int64)
    return out


# ----------------------------------------------------------------------------
# per-position / per-bin hotspot test
# ----------------------------------------------------------------------------

def nb_pvalue_exact(k, alpha, p):
    """nb_model.py:298-314 -- the reference's branches, evaluated with SciPy (scalar)."""
    import scipy.special
    import scipy.stats
    mu = alpha * (1 - p) / p
    if k < mu:
        return float(scipy.special.betainc(alpha, k + 1, p))
    pval = float(scipy.special.betainc(k, alpha, 1 - p))
    if pval == 0:
        pval = float(scipy.stats.nbinom.pmf(k, alpha, p))
    return pval


def base_probabilities(seq_chrom, start, end, s_prob, n_up=2, n_down=2, normed=True):
    """base_probabilities_by_region (sequence_tools.py:292-317) on an upper-cased chromosome byte array:
    returns (probs float64, positions int64).  s_prob is indexed by the k-mer's base-4 index."""
    L = len(seq_chrom)
    if start == 0:
        start = n_up                                            # fetch_sequence, sequence_tools.py:25-26
    f0, f1 = start - n_up, min(end + n_down, L)                 # faidx clips at the chromosome end
    f0 = min(f0, L)
    code = np.full(256, -1, dtype=np.int64)
    for i, b in enumerate(b"ACGT"):
        code[b] = i
        code[b | 0x20] = i
    c = code[np.asarray(seq_chrom[f0:f1], dtype=np.uint8)]
    n = len(c)
    poss = np.arange(f0 + n_up, f0 + max(n - n_down, n_up), dtype=np.int64)
    probs = np.zeros(len(poss), dtype=np.float64)
    klen = n_up + n_down + 1
    for j in range(len(poss)):
        w = c[j:j + klen]
        if (w < 0).any():
            continue
        idx = 0
        for v in w:
            idx = idx * 4 + int(v)
        probs[j] = s_prob[idx]
    if normed:
        with np.errstate(all="ignore"):
            probs = probs / np.sum(probs)
    return probs, poss


def position_test(seq_chrom, start, end, mu, sigma, s_prob, mut_starts, n_up=2, n_down=2, binsize=1):
    """apply_nb_to_region (nb_model.py:126-186): returns (pvals, poss, obss, exps, pts) for one region."""
    probs, pos_lst = base_probabilities(seq_chrom, start, end, s_prob, n_up, n_down, normed=True)
    vals, cnts = np.unique(np.asarray(mut_starts, dtype=np.int64), return_counts=True)
    mut_counts = dict(zip(vals.tolist(), cnts.tolist()))
    alpha, theta = mu ** 2 / sigma ** 2, sigma ** 2 / mu
    pvals, poss, obss, exps, pts = [], [], [], [], []
    for i in range(0, len(pos_lst), binsize):
        pt = probs[i] if binsize == 1 else np.sum(probs[i:i + binsize])
        k = sum(mut_counts.get(int(pos), 0) for pos in pos_lst[i:i + binsize])
        with np.errstate(all="ignore"):
            p = 1 / (pt * theta + 1)
            pvals.append(nb_pvalue_exact(k, alpha, p))
        poss.append(float(np.mean(pos_lst[i:i + binsize])))
        obss.append(k)
        exps.append(pt * mu)
        pts.append(pt)
    return (np.array(pvals, dtype=float), np.array(poss), np.array(obss), np.array(exps), np.array(pts))


# ----------------------------------------------------------------------------
# per-window observed counts: add_objectives, mutation branch
# ----------------------------------------------------------------------------

def muts_per_sample_per_element(df_mut, df_blocks, drop_duplicates=True):
    """tabulate_muts_per_sample_per_element (mutation_tools.py:191-230): rows (ELT, SAMPLE, OBS_SNV, OBS_INDEL,
    OBS_MUT).  Same restatement of ``bedtools intersect -wa -wb`` as tabulate_mutations_in_element above."""
    import pandas as pd
    hits = []
    for chrom, dm in df_mut.groupby("CHROM", sort=False):
        db = df_blocks[df_blocks.CHROM == chrom]
        if len(db) == 0:
            continue
        bs, be, elt = db.START.values, db.END.values, db.ELT.values
        for row in dm.itertuples(index=True):
            m = (bs < row.END) & (be > row.START)
            for e in elt[m]:
                hits.append((row.Index, e))
    if not hits:
        return pd.DataFrame({'ELT': [], 'SAMPLE': [], 'OBS_SNV': [], 'OBS_INDEL': [], 'OBS_MUT': []})
    h = pd.DataFrame(hits, columns=["ROW", "ELT"])
    df = df_mut.loc[h.ROW.values].reset_index(drop=True)
    df["ELT"] = h.ELT.values
    if drop_duplicates:
        df = df.drop_duplicates(["CHROM", "START", "END", "REF", "ALT", "SAMPLE", "ELT"])
    is_indel = df.ANNOT == "INDEL"
    snv = df[~is_indel].groupby(["ELT", "SAMPLE"]).size().reset_index(name="OBS_SNV")
    ind = df[is_indel].groupby(["ELT", "SAMPLE"]).size().reset_index(name="OBS_INDEL")
    cnt = snv.merge(ind, how="outer")
    cnt["OBS_SNV"] = cnt.OBS_SNV.fillna(0)
    cnt["OBS_INDEL"] = cnt.OBS_INDEL.fillna(0)
    cnt["OBS_MUT"] = cnt.OBS_SNV + cnt.OBS_INDEL
    return cnt[["ELT", "SAMPLE", "OBS_SNV", "OBS_INDEL", "OBS_MUT"]]


def window_objectives(df_mut, idx, max_muts_per_elt_per_sample=None, sample_filter_stdev=None,
                      max_muts_per_sample=None, filters=None):
    """add_objectives, mutation branch (DataExtractor.py:540-559).  ``filters`` may supply the reference's own
    (cap_muts_per_element_per_sample, filter_samples_by_stdev, filter_hypermut_samples) -- the golden generator
    passes the unmodified functions; by default the restatements of mutation_tools.py:293-326 below are used."""
    import pandas as pd

    def cap(df, m):                                              # :318-326 -- caps OBS_MUT only
        df.loc[df.OBS_MUT > m, 'OBS_MUT'] = m
        return df

    def by_stdev(df, cutoff):                                    # :306-316
        cnt = df.SAMPLE.value_counts()
        bl = cnt[cnt > cnt.std() * cutoff].index.to_list()
        return df[~df.SAMPLE.isin(bl)]

    def hypermut(df, m):                                         # :293-304
        cnt = df.SAMPLE.value_counts()
        bl = cnt[cnt > m].index.to_list()
        return df[~df.SAMPLE.isin(bl)]

    f_cap, f_std, f_hyp = filters or (cap, by_stdev, hypermut)
    idx = np.asarray(idx)
    df_idx = pd.DataFrame(idx, columns=['CHROM', 'START', 'END'])
    df_idx['ELT'] = ['{}:{}-{}'.format(r[0], r[1], r[2]) for r in idx]
    blocks = df_idx.copy()
    blocks['CHROM'] = blocks.CHROM.astype(str)
    df = muts_per_sample_per_element(df_mut, blocks, drop_duplicates=True)
    if max_muts_per_elt_per_sample:
        df = f_cap(df, max_muts_per_elt_per_sample)
    if sample_filter_stdev:
        df = f_std(df, sample_filter_stdev)
    if max_muts_per_sample:
        df = f_hyp(df, max_muts_per_sample)
    if len(df) == 0:
        return np.zeros(len(idx), dtype=np.int64)
    df_elt = df.pivot_table(index='ELT', values='OBS_SNV', aggfunc='sum')
    df_cnt = df_idx.merge(df_elt, on='ELT', how='left')
    df_cnt.loc[df_cnt.OBS_SNV.isna(), 'OBS_SNV'] = 0
    return df_cnt.OBS_SNV.astype(int).values.astype(np.int64)


# ----------------------------------------------------------------------------
# secondary gene tests
# ----------------------------------------------------------------------------

def gene_dnds_sel(alpha, theta, pi6, obs6):
    """gene_expected_muts_dnds, gene_pvalue_burden_dnds, gene_pvalue_sel_nb (transfer_tools.py:363-392, :617-676,
    :1172-1214, :1253-1276), vectorised with SciPy.  pi6 / obs6 columns: SYN, MIS, NONS, SPL, TRUNC, NONSYN.
    Returns a dict of [n] arrays with the reference's column names."""
    import scipy.stats
    alpha, theta = np.asarray(alpha, dtype=np.float64), np.asarray(theta, dtype=np.float64)
    pi6, obs6 = np.asarray(pi6, dtype=np.float64), np.asarray(obs6, dtype=np.float64)
    cls = ("SYN", "MIS", "NONS", "SPL", "TRUNC", "NONSYN")
    out = {}
    with np.errstate(all="ignore"):
        for j, c in enumerate(cls):
            out["EXP_" + c] = alpha * theta * pi6[:, j]
        th = theta * pi6[:, 0]
        tml = (obs6[:, 0] + alpha - 1) / (1 + (1 / th))
        lo = alpha * th
        tml = np.where(alpha <= 1, np.where(tml > lo, tml, lo), tml)             # Python max(lo, tml)
        out["T_SYN"] = tml
        ratio = tml / out["EXP_SYN"]
        out["MRFOLD"] = np.where(ratio > 1e-10, ratio, 1e-10)                    # Python max(1e-10, ratio)
        for j, c in enumerate(cls):
            out["EXP_%s_ML" % c] = out["EXP_" + c] * out["MRFOLD"]
            out["PVAL_%s_BURDEN_DNDS" % c] = nb_pvalue_greater_midp(obs6[:, j], alpha, 1 / (out["EXP_%s_ML" % c] / alpha + 1))
        ll = lambda k, t: scipy.stats.nbinom.logpmf(k, alpha, 1 / (1 + t))
        l0 = {j: ll(obs6[:, j], theta * pi6[:, j] * out["MRFOLD"]) for j in (0, 1, 4)}
        l1 = {j: ll(obs6[:, j], obs6[:, j] / alpha) for j in (0, 1, 4)}
        ll0 = l0[0] + l0[1] + l0[4]
        out["PVAL_SYN_SEL_NB"] = scipy.stats.chi2.sf(-2 * (ll0 - (l1[0] + l0[1] + l0[4])), df=1)
        out["PVAL_MIS_SEL_NB"] = scipy.stats.chi2.sf(-2 * (ll0 - (l0[0] + l1[1] + l0[4])), df=1)
        out["PVAL_TRUNC_SEL_NB"] = scipy.stats.chi2.sf(-2 * (ll0 - (l0[0] + l0[1] + l1[4])), df=1)
        out["PVAL_NONSYN_SEL_NB"] = scipy.stats.chi2.sf(-2 * (ll0 - (l0[0] + l1[1] + l1[4])), df=2)
    return out


def selection_coefficient(obs, exp, alpha, theta, pi):
    """selection_coefficient (transfer_tools.py:1279-1292): (SEL, PVAL_SEL)."""
    import scipy.stats
    with np.errstate(all="ignore"):
        sel = (obs + 1e-16) / (exp + 1e-16)
        ll0 = scipy.stats.nbinom.logpmf(obs, alpha, 1 / (1 + theta * pi))
        ll1 = scipy.stats.nbinom.logpmf(obs, alpha, 1 / (1 + theta * pi * sel))
        return sel, scipy.stats.chi2.sf(-2 * (ll0 - ll1), df=1)
