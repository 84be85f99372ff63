#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference); the GPU box sees just the
committed .npz files.  The reference ships no tests or known-answer vectors of its own
(SURVEY.md section 4), so these frozen outputs of its own functions are what pins the
oracle (oracle/dig_oracle.*) and, through it, the CUDA kernels.

    python tests/golden/make_golden.py

Every array saved here is either a seeded input or the direct return value of a
reference function; the glue that cannot be imported (h5py/bedtools-bound loops) is
re-typed from the cited lines and marked GLUE below.
"""
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402


def save(path, d):
    """np.savez without pickled object arrays (strings become fixed-width unicode)."""
    clean = {}
    for k, v in d.items():
        v = np.asarray(v)
        if v.dtype == object:
            v = v.astype(str)
        clean[k] = v
    np.savez_compressed(path, **clean)


def synth_seq(rng, n, n_runs=3, lower=True):
    s = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
    if lower:
        s = np.where(rng.random(n) < 0.5, s | 0x20, s).astype(np.uint8)
    for _ in range(n_runs):
        a = int(rng.integers(0, n - 50))
        b = a + int(rng.integers(1, 400))
        s[a:b] = ord("N")
    return s


def main():
    ref = rh.load_reference()
    st, gd, nb, tt, mt = (ref.sequence_tools, ref.genic_driver_tools, ref.nb_model,
                          ref.transfer_tools, ref.mutation_tools)
    rng = np.random.default_rng(20261017)

    # ---------------------------------------------------------------- scan
    seqs = {"chr1": synth_seq(rng, 40123), "chr2": synth_seq(rng, 25007)}
    seqs["chr2"][:7] = ord("N")            # N at the very start of a chromosome
    seqs["chr1"][-3:] = ord("N")
    rh.register_fasta("golden.fa", {k: v.tobytes().decode() for k, v in seqs.items()})
    W = 1000
    wins = []
    for ci, (name, s) in enumerate(seqs.items(), start=1):
        i = 0
        while i + W < len(s):               # DataExtractor.py:70-77 tiling rule
            wins.append((ci, i, i + W))
            i += W
        wins.append((ci, i, len(s)))        # a window touching the chromosome end (clip quirk)
        wins.append((ci, i, i + W))         # a window running past the chromosome end
    wins += [(1, 5, 6), (1, 17, 17), (1, 2, 40), (2, 3, 9), (1, 12345, 23456), (2, 1, 25007)]
    wins = np.array(wins, dtype=np.int64)
    scan = {"seq_chr1": seqs["chr1"], "seq_chr2": seqs["chr2"], "windows": wins}
    for (u, d) in [(1, 1), (2, 2), (0, 0), (1, 2), (2, 0)]:
        ok = wins[(wins[:, 1] == 0) | (wins[:, 1] >= u)]
        df = st.count_contexts_by_regions("golden.fa", ["chr%d" % c for c in ok[:, 0]], ok[:, 1], ok[:, 2],
                                          n_up=u, n_down=d)
        assert list(df.columns) == list(st.mk_context_sequences(u, d).keys())
        scan["counts_%d_%d" % (u, d)] = df.values.astype(np.int64)
        scan["rows_%d_%d" % (u, d)] = np.flatnonzero((wins[:, 1] == 0) | (wins[:, 1] >= u))
        scan["index_%d_%d" % (u, d)] = np.array(df.index)
    scan["columns_1_1"] = np.array(list(st.mk_context_sequences(1, 1).keys()))
    scan["columns_2_2"] = np.array(list(st.mk_context_sequences(2, 2).keys()))

    # count_contexts_in_bed through its own multiprocessing.Pool (sequence_tools.py:96-128)
    df_bed = pd.DataFrame(wins[:60], columns=[0, 1, 2])
    df_pool = st.count_contexts_in_bed("golden.fa", df_bed, n_up=1, n_down=1, N_proc=2, N_chunk=4)
    scan["pool_counts_1_1"] = df_pool.values.astype(np.int64)
    scan["pool_index"] = np.array(df_pool.index)

    # strand-aware block counts (nonc_elt_context_count, sequence_tools.py:527-566)
    nblk = 40
    bc = rng.integers(1, 3, nblk)
    bs = np.array([rng.integers(1, len(seqs["chr%d" % c]) - 700) for c in bc])
    be = bs + rng.integers(1, 600, nblk)
    bstr = rng.choice(np.array(["+", "-"]), nblk)
    regions = [(int(c), int(s), int(e), str(t)) for c, s, e, t in zip(bc, bs, be, bstr)]
    df192 = st.nonc_elt_context_count(regions, st.mk_trans_idx(), "golden.fa")
    scan["blk_chrom"], scan["blk_start"], scan["blk_end"] = bc, bs, be
    scan["blk_strand"] = np.where(bstr == "-", -1, 1).astype(np.int8)
    scan["blk_L192"] = df192.values.astype(np.float64)
    scan["blk_columns"] = np.array(df192.columns)
    save(os.path.join(HERE, "scan.npz"), scan)

    # ---------------------------------------------------------------- index tables
    tabs = {
        "trans_idx": np.array(st.mk_trans_idx()),
        "mutctx_mut": np.array(st.mk_mutation_context(return_df=True).MUT_TYPE),
        "mutctx_ctx": np.array(st.mk_mutation_context(return_df=True).CONTEXT),
    }
    # GLUE: the reference's minus-strand re-ordering expression (sequence_tools.py:610-614, :633-634)
    df_mc = st.mk_mutation_context(return_df=True)
    mut_model_idx = [r[1] + '>' + r[1][0] + r[0][2] + r[1][2] for r in zip(df_mc.MUT_TYPE, df_mc.CONTEXT)]
    subst_idx = sorted(mut_model_idx)
    revc_subst_idx = [st.reverse_complement(sub.split('>')[0]) + '>' + st.reverse_complement(sub.split('>')[-1])
                      for sub in subst_idx]
    revc_dic = dict(zip(subst_idx, revc_subst_idx))
    region_counts = np.arange(192) * 7 + 3
    flipped = [r[1] for r in sorted(enumerate(region_counts), key=lambda k: revc_dic[subst_idx[k[0]]])]
    tabs["revc_in"] = region_counts
    tabs["revc_out"] = np.array(flipped)
    save(os.path.join(HERE, "tables.npz"), tabs)

    # ---------------------------------------------------------------- mutation contexts
    muts = []
    for ci in (1, 2):
        s = seqs["chr%d" % ci]
        up = np.char.upper(s.tobytes().decode())
        pos = np.sort(rng.integers(0, len(s), 300))
        for p in pos:
            refb = str(up)[p]
            r = rng.random()
            if r < 0.1:
                refb = "ACGT"[(("ACGTN".index(refb)) + 1) % 4]      # REF mismatch
            elif r < 0.13:
                refb = refb + "A"                                   # multi-base REF
            alt = "ACGT"[int(rng.integers(0, 4))]
            muts.append((ci, int(p), int(p) + 1, refb, alt, "S%d" % rng.integers(0, 9), ".", "Noncoding"))
            if rng.random() < 0.25:                                 # same-START run (cache quirk)
                refb2 = str(up)[p] if rng.random() < 0.7 else "ACGT"[int(rng.integers(0, 4))]
                muts.append((ci, int(p), int(p) + 1, refb2, "ACGT"[int(rng.integers(0, 4))],
                             "S%d" % rng.integers(0, 9), ".", "Noncoding"))
        for p in (0, 1, 2, len(s) - 3, len(s) - 4):
            muts.append((ci, p, p + 1, str(up)[p], "A", "S0", ".", "Noncoding"))
    df_mut = pd.DataFrame(muts, columns=["CHROM", "START", "END", "REF", "ALT", "SAMPLE", "GENE", "ANNOT"])
    df_mut = df_mut[df_mut.REF != "N"].sort_values(["CHROM", "START"], kind="stable").reset_index(drop=True)
    mc = {}
    for (u, d) in [(1, 1), (2, 2)]:
        outs = []
        for chrom, g in df_mut.groupby("CHROM"):
            g = g.copy()
            g["ROW"] = g.index
            # drop rows whose context window would leave the chromosome (reference behaviour there is
            # an artefact of Python slicing: '' on the left edge, a truncated string on the right)
            L = len(seqs["chr%d" % chrom])
            o = st.mutation_contexts_by_chrom("golden.fa", g, n_up=u, n_down=d)
            o = o[(o.START - u >= 0) & (o.START + d < L)]
            outs.append(o)
        o = pd.concat(outs)
        mc["kept_rows_%d_%d" % (u, d)] = o.ROW.values.astype(np.int64)
        mc["context_%d_%d" % (u, d)] = np.array(o.CONTEXT)
        mc["mut_type_%d_%d" % (u, d)] = np.array(o.MUT_TYPE)
    for col in df_mut.columns:
        mc["in_" + col] = np.array(df_mut[col])
    save(os.path.join(HERE, "mutctx.npz"), mc)

    # ---------------------------------------------------------------- overlaps + region params
    ov = {}
    cases = [([15000], [29990]), ([20000], [30000]), ([20000], [20000]), ([19999], [20001]),
             ([0], [1]), ([5, 25000, 61000], [9000, 25100, 70000]), ([30000, 10], [40001, 20]),
             ([123456, 123999, 130000], [123789, 124100, 130000])]
    for i, (s, e) in enumerate(cases):
        for W2 in (10000, 1000):
            o = gd.get_ideal_overlaps(3, np.vstack((s, e)), W2)
            ov["case%d_w%d_in" % (i, W2)] = np.vstack((s, e))
            ov["case%d_w%d_out" % (i, W2)] = np.array(sorted(o), dtype=np.int64).reshape(-1, 3)
    save(os.path.join(HERE, "overlaps.npz"), ov)

    # ---------------------------------------------------------------- element transfer (DIG_onthefly loop)
    window = 1000
    idx_rows = [(c, s, e) for (c, s, e) in wins[:-6] if e - s == window and e <= len(seqs["chr%d" % c])]
    idx_rows = sorted(set(idx_rows))
    idx_arr = np.array(idx_rows, dtype=np.int64)
    nW = len(idx_arr)
    all_windows_df = pd.DataFrame({
        "CHROM": idx_arr[:, 0], "START": idx_arr[:, 1], "END": idx_arr[:, 2],
        "Y_TRUE": rng.poisson(20, nW), "Y_PRED": rng.gamma(2.0, 10.0, nW),
        "STD": rng.uniform(0.5, 5.0, nW), "FLAG": rng.random(nW) < 0.1,
    }, index=["chr%d:%d-%d" % tuple(r) for r in idx_arr])
    d_pr_192 = np.exp(rng.normal(np.log(1e-6), 1.0, 192))       # FREQ in mk_mutation_context order
    d_pr = pd.DataFrame(d_pr_192, mut_model_idx).sort_index()[0].values   # genic_driver_tools.py:324-325
    trans_idx = st.mk_trans_idx()
    nE = 30
    et = {"window": window, "idx": idx_arr, "Y_TRUE": all_windows_df.Y_TRUE.values,
          "Y_PRED": all_windows_df.Y_PRED.values, "STD": all_windows_df.STD.values,
          "FLAG": all_windows_df.FLAG.values, "FREQ_192": d_pr_192, "d_pr_sorted": d_pr}
    e_chrom, e_strand, e_ptr, e_bs, e_be = [], [], [0], [], []
    res = {k: [] for k in ("MU", "SIGMA", "R_OBS", "FLAG", "P_SUM", "R_SIZE", "ELT_SIZE", "P_INDEL", "ALPHA", "THETA")}
    L_all = []
    for i in range(nE):
        chrom = int(rng.integers(1, 3))
        Lc = (len(seqs["chr%d" % chrom]) // window) * window
        nb_ = int(rng.integers(1, 5))
        starts = np.sort(rng.integers(1, Lc - 900, nb_))
        ends = np.minimum(starts + rng.integers(1, 800, nb_), Lc - 1)
        strand = "-" if rng.random() < 0.5 else "+"
        # ---- GLUE: onthefly_tools.py:116-165 re-typed, every call is a reference function
        elts_as_intervals = np.vstack((starts, ends))
        overlaps = gd.get_ideal_overlaps(chrom, elts_as_intervals, window)
        chrom_lst, start_lst, end_lst = (['chr' + str(r[0]) for r in overlaps], [r[1] for r in overlaps],
                                         [r[2] for r in overlaps])
        region_df = st.count_contexts_by_regions("golden.fa", chrom_lst, start_lst, end_lst, n_up=1, n_down=1)
        region_counts = np.array([np.repeat(region, 3) for region in region_df.values]).sum(axis=0)
        if strand == '-1' or strand == '-':
            region_counts = np.array([r[1] for r in sorted(enumerate(region_counts),
                                                           key=lambda k: revc_dic[subst_idx[k[0]]])])
        L_ctx = st.nonc_elt_context_count([(chrom, int(s), int(e), strand) for s, e in zip(starts, ends)],
                                          trans_idx, "golden.fa")
        L_ctx = L_ctx.loc[~L_ctx.index.duplicated()]               # sequence_tools.py:525
        L = np.zeros((192))
        for s_, e_ in zip(starts, ends):
            L += L_ctx.loc['chr{}:{}-{}'.format(chrom, s_, e_)].values
        prob_sum = region_counts * d_pr
        t_pi = d_pr / prob_sum.sum()
        p_mut = (t_pi * L).sum()
        mu, sigma, R_obs, FLAG = gd.get_region_params_direct(all_windows_df, overlaps, window)
        alpha, theta = nb.normal_params_to_gamma(mu, sigma)
        # ---- end GLUE
        e_chrom.append(chrom)
        e_strand.append(-1 if strand == "-" else 1)
        e_bs += list(starts)
        e_be += list(ends)
        e_ptr.append(len(e_bs))
        L_all.append(L)
        for k, v in (("MU", mu), ("SIGMA", sigma), ("R_OBS", R_obs), ("FLAG", bool(FLAG)), ("P_SUM", p_mut),
                     ("R_SIZE", int(region_counts.sum() / 3)), ("ELT_SIZE", int(np.sum(L) / 3)),
                     ("ALPHA", alpha), ("THETA", theta)):
            res[k].append(v)
        res["P_INDEL"].append(res["ELT_SIZE"][-1] / res["R_SIZE"][-1])
    et.update(elt_chrom=np.array(e_chrom), elt_strand=np.array(e_strand, dtype=np.int8),
              blk_ptr=np.array(e_ptr), blk_start=np.array(e_bs), blk_end=np.array(e_be), L192=np.array(L_all))
    for k, v in res.items():
        et["out_" + k] = np.array(v)
    save(os.path.join(HERE, "transfer.npz"), et)

    # ---------------------------------------------------------------- NB p-values + test stage
    ks, als, ps = [], [], []
    for a in [1e-6, 1e-3, 0.05, 0.5, 1.0, 2.5, 5.0, 17.3, 100.0, 1234.5, 1e4, 1e6, 1e8]:
        for p in [1e-8, 1e-4, 0.01, 0.1, 0.3, 0.5, 0.7, 0.9, 0.99, 0.9999, 1 - 1e-9, 1.0]:
            for k in [0, 1, 2, 3, 5, 10, 30, 100, 300, 500, 1000, 10000]:
                ks.append(k), als.append(a), ps.append(p)
    # around-the-mean cases (k ~ mean +- few sd) where the continued fraction is slowest
    for a in [0.5, 5.0, 50.0, 500.0, 5e4, 5e6]:
        for p in [0.001, 0.05, 0.5, 0.95, 0.999]:
            mean = a * (1 - p) / p
            sd = np.sqrt(a * (1 - p)) / p
            for z in [-3, -1, -0.1, 0, 0.1, 1, 3, 6, 10]:
                k = np.floor(mean + z * sd)
                if 0 <= k < 1e7:
                    ks.append(k), als.append(a), ps.append(p)
    ks += [2.5, 0.5, 7.0, 3.0, 0.0]
    als += [5.0, 5.0, np.nan, 5.0, 0.0]
    ps += [0.5, 0.5, 0.5, np.nan, 0.5]
    ks, als, ps = np.array(ks, float), np.array(als, float), np.array(ps, float)
    with np.errstate(all="ignore"):
        pv = nb.nb_pvalue_greater_midp(ks, als, ps)
    nbg = {"k": ks, "alpha": als, "p": ps, "pval": pv}

    # realistic element table through the reference's own transfer_tools functions
    nE = 4000
    MU = rng.gamma(2.0, 20.0, nE)
    SIGMA = MU * rng.uniform(0.05, 0.5, nE)
    Pi = rng.uniform(1e-4, 0.3, nE)
    Pi_ind = rng.uniform(1e-4, 0.3, nE)
    cj, cj_indel = 1.37, 0.21
    alpha, theta = nb.normal_params_to_gamma(MU, SIGMA)
    df_pre = pd.DataFrame({"ELT_SIZE": 100, "FLAG": False, "R_SIZE": 10000, "R_OBS": 5, "R_INDEL": 5,
                           "MU": MU, "SIGMA": SIGMA, "ALPHA": alpha, "THETA": theta,
                           "MU_INDEL": MU, "SIGMA_INDEL": SIGMA, "ALPHA_INDEL": alpha, "THETA_INDEL": theta,
                           "Pi_SUM": Pi, "Pi_INDEL": Pi_ind}, index=["E%d" % i for i in range(nE)])
    obs_snv = rng.poisson(MU * cj * Pi)
    obs_snv[::50] += rng.integers(5, 60, len(obs_snv[::50]))      # some true "drivers"
    obs_samp = np.minimum(obs_snv, rng.binomial(obs_snv, 0.9))
    obs_ind = rng.poisson(MU * cj_indel * Pi_ind)
    df_tab = pd.DataFrame({"OBS_SAMPLES": obs_samp, "OBS_SNV": obs_snv, "OBS_INDEL": obs_ind},
                          index=df_pre.index)[: nE - 100]          # last 100 elements: no mutations (NaN -> 0)
    dfm = tt.transfer_element_model_with_indels(df_tab, df_pre, cj)
    dfm = tt.element_expected_muts_nb(dfm)
    dfm = tt.element_pvalue_burden_nb(dfm)
    dfm = tt.element_pvalue_burden_nb_by_sample(dfm)
    dfm = tt.element_pvalue_indel(dfm, cj_indel)
    import scipy.stats
    x2 = -2 * (np.log(dfm.PVAL_SNV_BURDEN) + np.log(dfm.PVAL_INDEL_BURDEN))   # transfer_tools.py:1086-1087
    dfm['PVAL_MUT_BURDEN'] = scipy.stats.chi2.sf(x2, df=4)
    nbg.update(elt_MU=MU, elt_SIGMA=SIGMA, elt_Pi_SUM=Pi, elt_Pi_INDEL=Pi_ind, elt_cj=cj, elt_cj_indel=cj_indel,
               elt_OBS_SNV=dfm.OBS_SNV.values, elt_OBS_SAMPLES=dfm.OBS_SAMPLES.values,
               elt_OBS_INDEL=dfm.OBS_INDEL.values)
    for col in ("ALPHA", "THETA", "THETA_INDEL", "EXP_SNV", "EXP_INDEL", "PVAL_SNV_BURDEN", "PVAL_SAMPLE_BURDEN",
                "PVAL_INDEL_BURDEN", "PVAL_MUT_BURDEN"):
        nbg["elt_out_" + col] = dfm[col].values.astype(np.float64)
    save(os.path.join(HERE, "nbtest.npz"), nbg)

    # ---------------------------------------------------------------- gene observed counts + gene test
    nG = 300
    genes = ["G%03d" % i for i in range(nG)] + ["TP53"]
    annots = np.array(["Synonymous", "Missense", "Nonsense", "Essential_Splice", "INDEL", "Stop_loss"])
    nM = 6000
    gm = pd.DataFrame({
        "CHROM": rng.integers(1, 23, nM), "START": rng.integers(1, 10 ** 6, nM),
        "REF": rng.choice(list("ACGT"), nM), "ALT": rng.choice(list("ACGT"), nM),
        "SAMPLE": ["S%02d" % s for s in rng.zipf(1.5, nM) % 40],
        "GENE": rng.choice(genes, nM), "ANNOT": rng.choice(annots, nM, p=[.25, .45, .05, .05, .15, .05]),
    })
    gm["END"] = gm.START + 1
    gm = gm[["CHROM", "START", "END", "REF", "ALT", "SAMPLE", "GENE", "ANNOT"]]
    gm = pd.concat([gm, gm.iloc[:200]]).reset_index(drop=True)     # duplicated rows
    gm = mt.filter_hypermut_samples(gm, 1500)
    gg = {"in_" + c: np.array(gm[c]) for c in gm.columns}
    for cap in (3e9, 2):
        cnt = mt.mutations_per_gene(gm, max_muts_per_gene_per_sample=cap)
        tag = "cap%d" % (0 if cap > 100 else cap)
        gg[tag + "_genes"] = np.array(cnt.index)
        for c in ("OBS_SYN", "OBS_MIS", "OBS_NONS", "OBS_SPL", "OBS_INDEL"):
            gg[tag + "_" + c] = cnt[c].values.astype(np.int64)
    cnt = mt.mutations_per_gene(gm)
    MUg = rng.gamma(2.0, 20.0, nG + 1)
    SIGg = MUg * rng.uniform(0.05, 0.5, nG + 1)
    a_, t_ = nb.normal_params_to_gamma(MUg, SIGg)
    pis = rng.uniform(1e-4, 0.05, (nG + 1, 4))
    pre = pd.DataFrame({"CHROM": 1, "GENE_LENGTH": 1000, "R_SIZE": 10000, "R_OBS": 5, "R_INDEL": 5,
                        "MU": MUg, "SIGMA": SIGg, "ALPHA": a_, "THETA": t_,
                        "MU_INDEL": MUg, "SIGMA_INDEL": SIGg, "ALPHA_INDEL": a_, "THETA_INDEL": t_, "FLAG": False,
                        "Pi_SYN": pis[:, 0], "Pi_MIS": pis[:, 1], "Pi_NONS": pis[:, 2], "Pi_SPL": pis[:, 3],
                        "Pi_TRUNC": pis[:, 2] + pis[:, 3], "Pi_NONSYN": pis[:, 1] + pis[:, 2] + pis[:, 3],
                        "Pi_INDEL": rng.uniform(1e-3, 0.1, nG + 1)}, index=genes)
    cjg = 0.0123
    dfg = tt.transfer_gene_model(gm, cnt, pre, cjg)
    dfg = tt.gene_expected_muts_nb(dfg)
    dfg = tt.gene_pvalue_burden_nb(dfg)
    dfg = tt.gene_pvalue_burden_nb_by_sample(dfg)
    dfg = tt.gene_pvalue_indel(dfg)
    x2 = -2 * (np.log(dfg.PVAL_TRUNC_BURDEN) + np.log(dfg.PVAL_INDEL_BURDEN))   # transfer_tools.py:860-861
    dfg['PVAL_MUT_BURDEN'] = scipy.stats.chi2.sf(x2, df=4)
    gg["pre_genes"] = np.array(genes)
    for c in pre.columns:
        gg["pre_" + c] = pre[c].values
    gg["cj"] = cjg
    for c in dfg.columns:
        if c.startswith(("OBS_", "N_SAMP_", "EXP_", "PVAL_", "THETA")):
            gg["out_" + c] = dfg[c].values.astype(np.float64)
    save(os.path.join(HERE, "genes.npz"), gg)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
