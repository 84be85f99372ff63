"""Pins the CPU oracle (oracle/) against outputs of the UNMODIFIED reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pandas as pd

from conftest import golden, golden_genome, assert_pvals_close


def test_context_column_order(oracle):
    z = golden("scan")
    assert oracle.context_names(1, 1) == list(z["columns_1_1"])
    assert oracle.context_names(2, 2) == list(z["columns_2_2"])


def test_scan_counts_bit_exact(oracle):
    z = golden("scan")
    seq, off, ln = golden_genome()
    wins = z["windows"]
    for (u, d) in [(1, 1), (2, 2), (0, 0), (1, 2), (2, 0)]:
        rows = z["rows_%d_%d" % (u, d)]
        w = wins[rows]
        got, n_other = oracle.count_regions(seq, off, ln, w[:, 0] - 1, w[:, 1], w[:, 2], n_up=u, n_down=d)
        assert n_other == 0
        assert np.array_equal(got, z["counts_%d_%d" % (u, d)]), (u, d)
        assert list(z["index_%d_%d" % (u, d)]) == ["chr%d:%d-%d" % tuple(r) for r in w]


def test_scan_python_port_matches_reference(oracle):
    z = golden("scan")
    wins = z["windows"][z["rows_2_2"]][:25]
    seqs = {"chr1": z["seq_chr1"].tobytes().decode(), "chr2": z["seq_chr2"].tobytes().decode()}
    got = oracle.py_count_regions(seqs, ["chr%d" % c for c in wins[:, 0]], wins[:, 1], wins[:, 2], 2, 2)
    assert np.array_equal(got, z["counts_2_2"][:25])


def test_scan_pool_chunking(oracle):
    z = golden("scan")
    seq, off, ln = golden_genome()
    w = z["windows"][:60]
    got, _ = oracle.count_regions(seq, off, ln, w[:, 0] - 1, w[:, 1], w[:, 2], n_up=1, n_down=1)
    assert np.array_equal(got, z["pool_counts_1_1"])


def test_scan_rejects_start_inside_halo(oracle):
    seq, off, ln = golden_genome()
    import pytest
    with pytest.raises(ValueError):
        oracle.count_regions(seq, off, ln, [0], [1], [50], n_up=2, n_down=2)


def test_block_counts_strand_aware(oracle):
    z = golden("scan")
    seq, off, ln = golden_genome()
    c64, _ = oracle.count_regions(seq, off, ln, z["blk_chrom"] - 1, z["blk_start"], z["blk_end"],
                                  n_up=1, n_down=1, strand=z["blk_strand"])
    assert list(z["blk_columns"]) == oracle.substitution_names()
    assert np.array_equal(np.repeat(c64, 3, axis=1).astype(np.float64), z["blk_L192"])


def test_index_tables(oracle):
    z = golden("tables")
    assert oracle.substitution_names() == list(z["trans_idx"])
    tab = oracle.mutation_context_table()
    assert [m for m, _ in tab] == list(z["mutctx_mut"])
    assert [c for _, c in tab] == list(z["mutctx_ctx"])
    perm, _ = oracle.revcomp_permutation_192()
    assert np.array_equal(z["revc_in"][perm], z["revc_out"])
    # the permutation is "new[name] = old[revcomp(name)]", an involution
    assert np.array_equal(perm[perm], np.arange(192))


def test_mutation_contexts(oracle):
    z = golden("mutctx")
    seq, off, ln = golden_genome()
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    ref = np.array([code.get(r, 255) for r in z["in_REF"]], dtype=np.uint8)
    for (u, d) in [(1, 1), (2, 2)]:
        ctx = oracle.mutation_contexts(seq, off, ln, z["in_CHROM"] - 1, z["in_START"], ref, n_up=u, n_down=d)
        kept = np.flatnonzero(ctx >= 0)
        assert np.array_equal(kept, z["kept_rows_%d_%d" % (u, d)])
        names = np.array(oracle.context_names(u, d))
        assert list(names[ctx[kept]]) == list(z["context_%d_%d" % (u, d)])
        mt = ["%s>%s" % (r, a) for r, a in zip(z["in_REF"][kept], z["in_ALT"][kept])]
        assert mt == list(z["mut_type_%d_%d" % (u, d)])


def test_ideal_overlaps(oracle):
    z = golden("overlaps")
    for i in range(8):
        for W in (10000, 1000):
            blocks = z["case%d_w%d_in" % (i, W)]
            want = z["case%d_w%d_out" % (i, W)]
            got = oracle.ideal_overlaps(blocks[0], blocks[1], W)
            assert got == list(want[:, 1])
            assert np.all(want[:, 2] - want[:, 1] == W)


def _transfer_inputs(z):
    idx = z["idx"]
    win_index = {(int(c), int(s)): i for i, (c, s, e) in enumerate(idx)}
    return idx, win_index


def test_element_transfer(oracle):
    z = golden("transfer")
    seq, off, ln = golden_genome()
    idx, win_index = _transfer_inputs(z)
    w64, _ = oracle.count_regions(seq, off, ln, idx[:, 0] - 1, idx[:, 1], idx[:, 2], n_up=1, n_down=1)
    assert np.array_equal(oracle.model192_to_dpr(z["FREQ_192"]), z["d_pr_sorted"])
    # L from the oracle's own strand-aware block counts
    nb = len(z["blk_start"])
    blk_elt = np.repeat(np.arange(len(z["elt_chrom"])), np.diff(z["blk_ptr"]))
    c64, _ = oracle.count_regions(seq, off, ln, z["elt_chrom"][blk_elt] - 1, z["blk_start"], z["blk_end"],
                                  n_up=1, n_down=1, strand=z["elt_strand"][blk_elt])
    L = np.zeros((len(z["elt_chrom"]), 192))
    np.add.at(L, blk_elt, np.repeat(c64, 3, axis=1))
    assert nb == len(blk_elt)
    assert np.array_equal(L, z["L192"])
    out = oracle.element_transfer(z["elt_chrom"], z["elt_strand"], z["blk_ptr"], z["blk_start"], z["blk_end"], L,
                                  int(z["window"]), win_index, w64, z["Y_PRED"], z["STD"], z["Y_TRUE"], z["FLAG"],
                                  z["d_pr_sorted"])
    for k in ("MU", "SIGMA", "P_SUM", "P_INDEL"):
        np.testing.assert_allclose(out[k], z["out_" + k], rtol=1e-12)
    for k in ("R_OBS", "R_SIZE", "ELT_SIZE"):
        assert np.array_equal(out[k], z["out_" + k])
    assert np.array_equal(out["FLAG"], z["out_FLAG"])
    a, t = oracle.normal_params_to_gamma(out["MU"], out["SIGMA"])
    np.testing.assert_allclose(a, z["out_ALPHA"], rtol=1e-12)
    np.testing.assert_allclose(t, z["out_THETA"], rtol=1e-12)


def test_nb_midp_is_the_reference_expression(oracle):
    z = golden("nbtest")
    with np.errstate(all="ignore"):
        got = oracle.nb_pvalue_greater_midp(z["k"], z["alpha"], z["p"])
    assert_pvals_close(got, z["pval"], tol=1e-12)


def test_element_test_stage(oracle):
    z = golden("nbtest")
    alpha, theta = oracle.normal_params_to_gamma(z["elt_MU"], z["elt_SIGMA"])
    assert np.array_equal(alpha, z["elt_out_ALPHA"])
    theta_c = theta * float(z["elt_cj"])
    np.testing.assert_allclose(theta_c, z["elt_out_THETA"], rtol=0, atol=0)
    exp, p = oracle.burden_test(z["elt_OBS_SNV"], alpha, theta_c, z["elt_Pi_SUM"])
    np.testing.assert_allclose(exp, z["elt_out_EXP_SNV"], rtol=1e-15)
    assert_pvals_close(p, z["elt_out_PVAL_SNV_BURDEN"], tol=1e-12)
    _, ps = oracle.burden_test(z["elt_OBS_SAMPLES"], alpha, theta_c, z["elt_Pi_SUM"])
    assert_pvals_close(ps, z["elt_out_PVAL_SAMPLE_BURDEN"], tol=1e-12)
    theta_i = theta * float(z["elt_cj_indel"])
    exp_i, pi = oracle.burden_test(z["elt_OBS_INDEL"], alpha, theta_i, z["elt_Pi_INDEL"])
    np.testing.assert_allclose(exp_i, z["elt_out_EXP_INDEL"], rtol=1e-15)
    assert_pvals_close(pi, z["elt_out_PVAL_INDEL_BURDEN"], tol=1e-12)
    assert_pvals_close(oracle.fisher2(p, pi), z["elt_out_PVAL_MUT_BURDEN"], tol=1e-12)


def test_gene_observed_counts(oracle):
    z = golden("genes")
    df = pd.DataFrame({c: z["in_" + c] for c in ("CHROM", "START", "END", "REF", "ALT", "SAMPLE", "GENE", "ANNOT")})
    for cap, tag in ((3e9, "cap0"), (2, "cap2")):
        got = oracle.gene_observed_counts(df, max_muts_per_gene_per_sample=cap)
        assert list(got.index) == list(z[tag + "_genes"])
        for c in ("OBS_SYN", "OBS_MIS", "OBS_NONS", "OBS_SPL", "OBS_INDEL"):
            assert np.array_equal(got[c].values, z[tag + "_" + c]), (tag, c)
    got = oracle.gene_observed_counts(df).reindex(z["pre_genes"]).fillna(0)
    for c in ("N_SAMP_SYN", "N_SAMP_MIS", "N_SAMP_NONS", "N_SAMP_SPL", "N_SAMP_TRUNC", "N_SAMP_NONSYN",
              "N_SAMP_INDEL", "OBS_SYN", "OBS_MIS", "OBS_NONS", "OBS_SPL", "OBS_INDEL"):
        assert np.array_equal(got[c].values.astype(np.float64), z["out_" + c]), c


def test_tabulate_mutations_in_element_small(oracle):
    # hand-checked case for the bedtools-intersect restatement (mutation_tools.py:191-230, :155-189)
    mut = pd.DataFrame([
        (1, 100, 101, "A", "C", "S1", "Missense"),
        (1, 100, 101, "A", "C", "S1", "Missense"),     # exact duplicate row -> counted once
        (1, 150, 151, "G", "T", "S1", "Noncoding"),
        (1, 199, 200, "G", "T", "S2", "Noncoding"),    # last base of block [100,200)
        (1, 200, 201, "G", "T", "S2", "Noncoding"),    # first base past the block -> no hit
        (1, 120, 125, "GTTTT", "G", "S3", "INDEL"),
        (1, 298, 303, "GTTTT", "G", "S3", "INDEL"),    # spans blocks [250,300) and [300,350) of E2 -> once
        (2, 100, 101, "A", "C", "S1", "Missense"),     # other chromosome
    ], columns=["CHROM", "START", "END", "REF", "ALT", "SAMPLE", "ANNOT"])
    blocks = pd.DataFrame([(1, 100, 200, "E1"), (1, 250, 300, "E2"), (1, 300, 350, "E2"), (1, 110, 130, "E3")],
                          columns=["CHROM", "START", "END", "ELT"])
    summ, black = oracle.tabulate_mutations_in_element(mut, blocks)
    assert black == []
    assert summ.loc["E1"].tolist() == [3, 3, 1]
    assert summ.loc["E2"].tolist() == [1, 0, 1]
    assert summ.loc["E3"].tolist() == [1, 0, 1]
    summ, black = oracle.tabulate_mutations_in_element(mut, blocks, max_muts_per_sample=2.5)
    assert sorted(black) == ["S3"]
    assert summ.loc["E1"].tolist() == [2, 3, 0]
    summ, _ = oracle.tabulate_mutations_in_element(mut, blocks, max_muts_per_elt_per_sample=1)
    assert summ.loc["E1"].tolist() == [3, 2, 1]


def test_position_test_oracle_matches_reference_golden(oracle):
    """a16: the oracle's apply_nb_to_region / nb_pvalue_exact restatement against the frozen outputs of the
    unmodified reference (tests/golden/make_golden_position.py)."""
    z = golden("position")
    seq, muts = z["seq"], z["mut_start"]
    for ci, (u, d, binsize, s, e, mu, sigma, n) in enumerate(z["cases"]):
        u, d, binsize, s, e = int(u), int(d), int(binsize), int(s), int(e)
        pv, pos, obs, exp, pt = oracle.position_test(seq, s, e, mu, sigma, z["s_prob_%d_%d" % (u, d)], muts, u, d, binsize)
        assert len(pv) == int(n)
        assert np.array_equal(obs, z["obs_%d" % ci]) and np.array_equal(pos, z["pos_%d" % ci])
        assert np.array_equal(pt, z["pt_%d" % ci], equal_nan=True) and np.array_equal(exp, z["exp_%d" % ci], equal_nan=True)
        assert np.array_equal(pv, z["pval_%d" % ci], equal_nan=True)
    with np.errstate(all="ignore"):
        ex = np.array([oracle.nb_pvalue_exact(k, a, p) for k, a, p in zip(z["ex_k"], z["ex_alpha"], z["ex_p"])])
    assert np.array_equal(ex, z["ex_pval"], equal_nan=True)


def _objectives_frame(z):
    return pd.DataFrame({k: z["mut_" + k] for k in ("CHROM", "START", "END", "REF", "ALT", "SAMPLE", "GENE", "ANNOT")})


def test_window_objectives_oracle_matches_reference_filters(oracle):
    """f-2: the oracle's own restatement of the three sample filters against the goldens produced with the
    reference's unmodified filter functions."""
    z = golden("objectives")
    df = _objectives_frame(z)
    for i, c in enumerate(z["cases"]):
        cap, std, mx = [None if np.isnan(v) else v for v in c]
        got = oracle.window_objectives(df.copy(), z["idx"], cap, std, mx)
        assert np.array_equal(got, z["y_%d" % i]), i
    assert z["y_0"].sum() > z["y_3"].sum() > 0 and np.array_equal(z["y_0"], z["y_1"])   # the cap quirk: no effect


def _secondary_inputs(z):
    cls = ("SYN", "MIS", "NONS", "SPL", "TRUNC", "NONSYN")
    pi6 = np.stack([z["in_Pi_" + c] for c in cls], axis=1)
    obs6 = np.stack([z["in_OBS_" + c] for c in cls], axis=1)
    return z["in_ALPHA"], z["in_THETA"], pi6, obs6


def test_secondary_gene_tests_oracle_matches_reference_golden(oracle):
    """f-4: dN/dS expectations, dN/dS burden p-values, LLR selection tests and selection coefficients."""
    z = golden("secondary")
    alpha, theta, pi6, obs6 = _secondary_inputs(z)
    got = oracle.gene_dnds_sel(alpha, theta, pi6, obs6)
    for k, v in got.items():
        want = z["out_" + k]
        if k.startswith("PVAL"):
            assert_pvals_close(v, want, tol=1e-9)
        else:
            assert np.array_equal(v, want, equal_nan=True), k
    for j, c in ((0, "SYN"), (1, "MIS"), (4, "TRUNC")):
        sel, pv = oracle.selection_coefficient(obs6[:, j], z["out_EXP_" + c], alpha, theta, pi6[:, j])
        assert np.array_equal(sel, z["out_SEL_" + c], equal_nan=True)
        assert_pvals_close(pv, z["out_PVAL_%s_SEL" % c], tol=1e-9)


def test_check_all_windows_detects_one_wrong_bin(oracle):
    """oracle/parity_sample.check_all_windows (the exhaustive full-size check of the GPU suite and of bench.py): accepts the
    oracle's own rows chromosome by chromosome, sub-ranges included, and names the window when one bin is off by one."""
    from oracle import parity_sample as ps
    lengths = np.array([250_000, 130_005], dtype=np.int64)
    off = np.array([0, 250_112], dtype=np.int64)
    wins = np.array([(c, s, min(s + 10_000, int(L))) for c, L in enumerate(lengths) for s in range(0, int(L), 10_000)],
                    dtype=np.int64)
    rows = {2: [], 1: []}
    for c, L in enumerate(lengths):
        seq = oracle.synth_genome(int(off[c]), int(L), 3)
        sel = np.flatnonzero(wins[:, 0] == c)
        z = np.zeros(len(sel), dtype=np.int32)
        for nu in (2, 1):
            rows[nu].append(oracle.count_regions(seq, np.zeros(1, dtype=np.int64), np.array([L]), z, wins[sel, 1],
                                                 wins[sel, 2], nu, nu)[0])
    f5, f3 = np.concatenate(rows[2]), np.concatenate(rows[1])
    ok = ps.check_all_windows(lengths, off, 3, wins, lambda a, b: f5[a:b], lambda a, b: f3[a:b], f5.sum(0), f3.sum(0))
    assert ok == {"windows": len(wins), "ok": True, "detail": ""}
    part = ps.check_all_windows(lengths, off, 3, wins, lambda a, b: f5[a:b], lambda a, b: f3[a:b], lo=10, hi=30)
    assert part["ok"] and part["windows"] == 20
    f5[17, 3] += 1
    bad = ps.check_all_windows(lengths, off, 3, wins, lambda a, b: f5[a:b], lambda a, b: f3[a:b], f5.sum(0), f3.sum(0))
    assert not bad["ok"] and "first window 17" in bad["detail"] and "pentanucleotide totals differ" in bad["detail"]
