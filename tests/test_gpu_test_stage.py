"""GPU parity tests for K3 (mutation contexts), K5 (observed counts), K6 (element transfer) and
K7 (NB burden test), through the C ABI, against the golden vectors of the unmodified reference and
the CPU oracle.  Tolerances are the ones north_star states: counts bit-exact, expectations
rel. error <= 1e-9, p-values |dlog10 p| <= 1e-6."""
import numpy as np
import pandas as pd
import pytest
import torch

from conftest import golden, golden_genome, assert_pvals_close

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def gold_dev_genome():
    from digdriver_b200.genome import Genome, DeviceGenome
    z = golden("scan")
    return DeviceGenome.from_genome(Genome(["chr1", "chr2"], [z["seq_chr1"], z["seq_chr2"]]), DEV)


# ------------------------------------------------------------------ K3

def test_mutation_contexts_golden(gold_dev_genome, oracle):
    from digdriver_b200 import kernels
    z = golden("mutctx")
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    ref = np.array([code.get(r, 255) for r in z["in_REF"]], dtype=np.uint8)
    seq, off, ln = golden_genome()
    for (u, d) in [(1, 1), (2, 2)]:
        ctx = kernels.mutation_contexts(gold_dev_genome, z["in_CHROM"] - 1, z["in_START"], ref, u, d).cpu().numpy()
        kept = np.flatnonzero(ctx >= 0)
        assert np.array_equal(kept, z["kept_rows_%d_%d" % (u, d)])
        names = np.array(oracle.context_names(u, d))
        assert list(names[ctx[kept]]) == list(z["context_%d_%d" % (u, d)])
        want = oracle.mutation_contexts(seq, off, ln, z["in_CHROM"] - 1, z["in_START"], ref, u, d)
        assert np.array_equal(ctx, want)


def test_mutation_contexts_random_vs_oracle(oracle):
    from digdriver_b200 import kernels
    from digdriver_b200.genome import DeviceGenome
    lengths = np.array([400_000, 250_000], dtype=np.int64)
    dg = DeviceGenome.synthetic(["chr1", "chr2"], lengths, seed=11, device=DEV)
    seq = np.full(dg.n_bases, ord("N"), dtype=np.uint8)
    for o, n in zip(dg.chrom_off, lengths):
        seq[o:o + n] = oracle.synth_genome(int(o), int(n), 11)
    rng = np.random.default_rng(3)
    n = 50_000
    chrom = np.sort(rng.integers(0, 2, n)).astype(np.int32)
    start = rng.integers(0, lengths[chrom])
    order = np.lexsort((start, chrom))
    chrom, start = chrom[order], start[order]
    start[1::7] = start[0:-1:7][:len(start[1::7])]          # same-START runs
    up = seq & 0xDF
    true_ref = np.select([up[dg.chrom_off[chrom] + start] == ord(c) for c in "ACGT"], [0, 1, 2, 3], 255)
    ref = np.where(rng.random(n) < 0.15, rng.integers(0, 4, n), true_ref).astype(np.uint8)
    ref[::101] = 255
    for (u, d) in [(1, 1), (2, 2), (0, 3)]:
        got = kernels.mutation_contexts(dg, chrom, start, ref, u, d).cpu().numpy()
        want = oracle.mutation_contexts(seq, dg.chrom_off, lengths, chrom, start, ref, u, d)
        assert np.array_equal(got, want)


# ------------------------------------------------------------------ K7

def test_nb_midp_golden_grid():
    from digdriver_b200 import kernels
    z = golden("nbtest")
    got = kernels.nb_pvalue_greater_midp(z["k"], z["alpha"], z["p"], DEV).cpu().numpy()
    assert_pvals_close(got, z["pval"], tol=1e-6)


def test_nb_midp_edge_cases_match_scipy(oracle):
    from digdriver_b200 import kernels
    k = np.array([0, 1, 5, 0, 3, 2.5, 0, 7, np.nan, 4, 4, 300, 500, 0, 1e4])
    a = np.array([5, 5, 5, 5, 5, 5, np.nan, 5, 5, 1e8, 1e-6, 5, 5, 2.0, 1e8])
    p = np.array([1.0, 1.0, 1.0, 0.5, np.nan, 0.5, 0.5, 0.999999999999, 0.5, 0.99999, 0.5, 0.9, 0.9, 1e-300, 0.9999])
    with np.errstate(all="ignore"):
        want = oracle.nb_pvalue_greater_midp(k, a, p)
    got = kernels.nb_pvalue_greater_midp(k, a, p, DEV).cpu().numpy()
    assert_pvals_close(got, want, tol=1e-6)
    assert got[0] == 0.5 and got[1] == 0.0 and got[2] == 0.0       # Pi = 0: all mass at zero
    assert got[12] == 0.0 and want[12] == 0.0                      # deep tail underflows to exactly 0.0


def test_element_burden_stage_golden():
    """element_expected_muts_nb + element_pvalue_burden_nb(_by_sample) + element_pvalue_indel + Fisher."""
    from digdriver_b200 import kernels
    z = golden("nbtest")
    alpha, theta = z["elt_out_ALPHA"], z["elt_out_THETA"]
    exp, p = kernels.nb_burden_test(z["elt_OBS_SNV"], alpha, theta, z["elt_Pi_SUM"], DEV)
    assert np.array_equal(exp.cpu().numpy(), z["elt_out_EXP_SNV"])          # bit-exact expectations
    assert_pvals_close(p.cpu().numpy(), z["elt_out_PVAL_SNV_BURDEN"])
    _, ps = kernels.nb_burden_test(z["elt_OBS_SAMPLES"], alpha, theta, z["elt_Pi_SUM"], DEV, want_exp=False)
    assert_pvals_close(ps.cpu().numpy(), z["elt_out_PVAL_SAMPLE_BURDEN"])
    exp_i, pi = kernels.nb_burden_test(z["elt_OBS_INDEL"], alpha, z["elt_out_THETA_INDEL"], z["elt_Pi_INDEL"], DEV)
    np.testing.assert_allclose(exp_i.cpu().numpy(), z["elt_out_EXP_INDEL"], rtol=1e-15)
    assert_pvals_close(pi.cpu().numpy(), z["elt_out_PVAL_INDEL_BURDEN"])
    comb = kernels.fisher_combine2(p, pi, DEV).cpu().numpy()
    assert_pvals_close(comb, z["elt_out_PVAL_MUT_BURDEN"])


def test_gene_burden_stage_golden():
    from digdriver_b200 import kernels
    z = golden("genes")
    alpha, theta = z["pre_ALPHA"], z["out_THETA"]
    for cls, pi in (("SYN", "Pi_SYN"), ("MIS", "Pi_MIS"), ("NONS", "Pi_NONS"), ("SPL", "Pi_SPL"),
                    ("TRUNC", "Pi_TRUNC"), ("NONSYN", "Pi_NONSYN")):
        exp, p = kernels.nb_burden_test(z["out_OBS_" + cls], alpha, theta, z["pre_" + pi], DEV)
        np.testing.assert_allclose(exp.cpu().numpy(), z["out_EXP_" + cls], rtol=1e-15)
        assert_pvals_close(p.cpu().numpy(), z["out_PVAL_%s_BURDEN" % cls])
        _, ps = kernels.nb_burden_test(z["out_N_SAMP_" + cls], alpha, theta, z["pre_" + pi], DEV, want_exp=False)
        assert_pvals_close(ps.cpu().numpy(), z["out_PVAL_%s_BURDEN_SAMPLE" % cls])
    exp_i, pi_ = kernels.nb_burden_test(z["out_OBS_INDEL"], z["pre_ALPHA_INDEL"], z["out_THETA_INDEL"],
                                        z["pre_Pi_INDEL"], DEV)
    np.testing.assert_allclose(exp_i.cpu().numpy(), z["out_EXP_INDEL"], rtol=1e-15)
    assert_pvals_close(pi_.cpu().numpy(), z["out_PVAL_INDEL_BURDEN"])
    comb = kernels.fisher_combine2(z["out_PVAL_TRUNC_BURDEN"], z["out_PVAL_INDEL_BURDEN"], DEV).cpu().numpy()
    assert_pvals_close(comb, z["out_PVAL_MUT_BURDEN"])


def test_nb_midp_random_sweep_vs_scipy(oracle):
    """1M random (k, alpha, theta, Pi) rows in the regime of BASELINE.md's probe."""
    from digdriver_b200 import kernels
    rng = np.random.default_rng(8)
    n = 1_000_000
    mu = rng.gamma(2.0, 20.0, n)
    sigma = mu * rng.uniform(0.05, 0.5, n)
    pi = rng.uniform(1e-4, 0.3, n)
    alpha, theta = oracle.normal_params_to_gamma(mu, sigma)
    k = rng.poisson(mu * pi).astype(np.float64)
    k[::97] += rng.integers(1, 200, len(k[::97]))
    want_exp, want_p = oracle.burden_test(k, alpha, theta, pi)
    exp, p = kernels.nb_burden_test(k, alpha, theta, pi, DEV)
    assert np.array_equal(exp.cpu().numpy(), want_exp)
    assert_pvals_close(p.cpu().numpy(), want_p)


# ------------------------------------------------------------------ K5

def _keyed(chrom, pos):
    return (np.asarray(chrom, dtype=np.int64) << 32) | np.asarray(pos, dtype=np.int64)


def test_tabulate_elements_vs_oracle(oracle):
    from digdriver_b200 import kernels
    rng = np.random.default_rng(12)
    n_elt, n_sample, n_mut = 300, 40, 20000
    # overlapping multi-block elements on 3 chromosomes
    rows = []
    for e in range(n_elt):
        c = int(rng.integers(1, 4))
        s = int(rng.integers(0, 200_000))
        for _ in range(int(rng.integers(1, 5))):
            ln = int(rng.integers(1, 400))
            rows.append((c, s, s + ln, "E%d" % e))
            s += int(rng.integers(-50, 600))
            s = max(s, 0)
    blocks = pd.DataFrame(rows, columns=["CHROM", "START", "END", "ELT"])
    mc = rng.integers(1, 5, n_mut)
    ms = rng.integers(0, 210_000, n_mut)
    is_indel = rng.random(n_mut) < 0.15
    me = ms + np.where(is_indel, rng.integers(1, 30, n_mut), 1)
    mut = pd.DataFrame({"CHROM": mc, "START": ms, "END": me,
                        "REF": rng.choice(list("ACGT"), n_mut), "ALT": rng.choice(list("ACGT"), n_mut),
                        "SAMPLE": ["S%d" % s for s in rng.zipf(1.3, n_mut) % n_sample],
                        "ANNOT": np.where(is_indel, "INDEL", "Noncoding")})
    mut = pd.concat([mut, mut.iloc[:500]]).reset_index(drop=True)        # duplicated rows
    for mps, mpe in ((1e9, 3e9), (60, 3e9), (1e9, 2)):
        want, black = oracle.tabulate_mutations_in_element(mut, blocks, max_muts_per_sample=mps,
                                                           max_muts_per_elt_per_sample=mpe)
        dedup = mut.drop_duplicates(["CHROM", "START", "END", "REF", "ALT", "SAMPLE"])
        samples, sample_id = np.unique(dedup.SAMPLE.values, return_inverse=True)
        elts = np.array(["E%d" % e for e in range(n_elt)])
        elt_id = pd.Series(np.arange(n_elt), index=elts)[blocks.ELT.values].values
        obs, stot = kernels.tabulate_elements(
            _keyed(blocks.CHROM, blocks.START), _keyed(blocks.CHROM, blocks.END), elt_id,
            _keyed(dedup.CHROM, dedup.START), _keyed(dedup.CHROM, dedup.END), sample_id,
            (dedup.ANNOT.values == "INDEL"), n_elt, len(samples), max_muts_per_sample=int(mps),
            max_per_elt_per_sample=int(mpe), device=DEV)
        obs = obs.cpu().numpy()
        got = pd.DataFrame(obs, index=elts, columns=["OBS_SAMPLES", "OBS_SNV", "OBS_INDEL"])
        got = got[got.OBS_SAMPLES > 0]
        want = want.astype(np.int64).sort_index()
        got = got.sort_index()
        assert list(got.index) == list(want.index), (mps, mpe)
        assert np.array_equal(got.values, want[["OBS_SAMPLES", "OBS_SNV", "OBS_INDEL"]].values), (mps, mpe)
        assert sorted(samples[stot.cpu().numpy() > mps]) == sorted(black)


def test_tabulate_genes_golden():
    from digdriver_b200 import kernels
    z = golden("genes")
    genes = z["pre_genes"]
    gid = pd.Series(np.arange(len(genes)), index=genes)
    cls_map = {"Synonymous": 0, "Missense": 1, "Nonsense": 2, "Essential_Splice": 3, "INDEL": 4}
    samples, sid = np.unique(z["in_SAMPLE"], return_inverse=True)
    mg = gid.reindex(z["in_GENE"]).fillna(-1).values.astype(np.int32)
    mcls = np.array([cls_map.get(a, 255) for a in z["in_ANNOT"]], dtype=np.uint8)
    obs, nsamp = kernels.tabulate_genes(mg, sid, mcls, len(genes), device=DEV)
    obs, nsamp = obs.cpu().numpy(), nsamp.cpu().numpy()
    for j, c in enumerate(("OBS_SYN", "OBS_MIS", "OBS_NONS", "OBS_SPL", "OBS_INDEL")):
        assert np.array_equal(obs[:, j].astype(np.float64), z["out_" + c]), c
    for j, c in enumerate(("N_SAMP_SYN", "N_SAMP_MIS", "N_SAMP_NONS", "N_SAMP_SPL", "N_SAMP_TRUNC",
                           "N_SAMP_NONSYN", "N_SAMP_INDEL")):
        assert np.array_equal(nsamp[:, j].astype(np.float64), z["out_" + c]), c
    # per-(gene, sample, class) cap (mutation_tools.py:334)
    obs2, _ = kernels.tabulate_genes(mg, sid, mcls, len(genes), max_per_gene_per_sample=2, device=DEV)
    want = pd.DataFrame({c: z["cap2_" + c] for c in ("OBS_SYN", "OBS_MIS", "OBS_NONS", "OBS_SPL", "OBS_INDEL")},
                        index=z["cap2_genes"]).reindex(genes).fillna(0).values
    assert np.array_equal(obs2.cpu().numpy(), want.astype(np.int64))


# ------------------------------------------------------------------ K6

def test_element_transfer_golden(gold_dev_genome):
    from digdriver_b200 import kernels
    z = golden("transfer")
    idx = z["idx"]
    window = int(z["window"])
    w64, _ = kernels.count_contexts(gold_dev_genome, idx[:, 0] - 1, idx[:, 1], idx[:, 2], 1, 1)
    blk_elt = np.repeat(np.arange(len(z["elt_chrom"])), np.diff(z["blk_ptr"]))
    bc, _ = kernels.count_contexts(gold_dev_genome, z["elt_chrom"][blk_elt] - 1, z["blk_start"], z["blk_end"], 1, 1,
                                   strand=z["elt_strand"][blk_elt])
    off, wmap = kernels.build_window_map(idx[:, 0] - 1, idx[:, 1], window, 2)
    out = kernels.element_transfer(z["elt_chrom"] - 1, z["elt_strand"], z["blk_ptr"], z["blk_start"], z["blk_end"],
                                   window, off, wmap, w64, z["Y_PRED"], z["STD"], z["Y_TRUE"].astype(np.float64),
                                   z["FLAG"], z["d_pr_sorted"], blk_counts=bc, device=DEV)
    o = {k: v.cpu().numpy() for k, v in out.items()}
    np.testing.assert_allclose(o["MU"][0], z["out_MU"], rtol=1e-12)
    np.testing.assert_allclose(o["SIGMA"][0], z["out_SIGMA"], rtol=1e-12)
    np.testing.assert_allclose(o["P"][0, :, 0], z["out_P_SUM"], rtol=1e-9)
    assert np.array_equal(o["R_OBS"][0], z["out_R_OBS"].astype(np.float64))
    assert np.array_equal(o["FLAG"][0].astype(bool), z["out_FLAG"])
    assert np.array_equal(o["R_SIZE"], z["out_R_SIZE"])
    assert np.array_equal(o["ELT_SIZE"], z["out_ELT_SIZE"])
    np.testing.assert_allclose(o["ELT_SIZE"] / o["R_SIZE"], z["out_P_INDEL"], rtol=1e-15)


def test_element_transfer_missing_window_raises(gold_dev_genome):
    from digdriver_b200 import kernels
    z = golden("transfer")
    idx = z["idx"][5:]                                   # drop the first windows of chr1
    window = int(z["window"])
    w64, _ = kernels.count_contexts(gold_dev_genome, idx[:, 0] - 1, idx[:, 1], idx[:, 2], 1, 1)
    off, wmap = kernels.build_window_map(idx[:, 0] - 1, idx[:, 1], window, 2)
    n = len(idx)
    with pytest.raises(KeyError):
        kernels.element_transfer([0], [1], [0, 1], [100], [200], window, off, wmap, w64, np.ones(n), np.ones(n),
                                 np.ones(n), np.zeros(n), z["d_pr_sorted"], L_elt=np.ones((1, 192, 1)), device=DEV)


def test_gene_transfer_multi_cohort_vs_oracle(gold_dev_genome, oracle):
    """Gene mode (L[192,4] input, inclusive CDS intervals) with 3 cohorts in one launch."""
    from digdriver_b200 import kernels
    z = golden("transfer")
    rng = np.random.default_rng(77)
    idx = z["idx"]
    window = int(z["window"])
    nW = len(idx)
    w64, _ = kernels.count_contexts(gold_dev_genome, idx[:, 0] - 1, idx[:, 1], idx[:, 2], 1, 1)
    w64h = w64.cpu().numpy().astype(np.int64)
    win_index = {(int(c), int(s)): i for i, (c, s, e) in enumerate(idx)}
    E, C = 64, 3
    chrom = rng.integers(1, 3, E)
    strand = np.where(rng.random(E) < 0.5, -1, 1).astype(np.int8)
    ptr = [0]
    bs, be = [], []
    for i in range(E):
        Lc = 39000 if chrom[i] == 1 else 24000
        nb = int(rng.integers(1, 12))
        s = np.sort(rng.integers(1, Lc - 500, nb))
        if rng.random() < 0.3:
            s = s[::-1].copy()                           # unsorted blocks
        bs += list(s)
        be += list(s + rng.integers(0, 400, nb))
        ptr.append(len(bs))
    L = rng.integers(0, 50, (E, 192, 4)).astype(np.float64)
    yp = rng.gamma(2.0, 10.0, (C, nW))
    sd = rng.uniform(0.5, 5.0, (C, nW))
    yt = rng.poisson(20, (C, nW)).astype(np.float64)
    fl = rng.random((C, nW)) < 0.1
    dpr = np.exp(rng.normal(np.log(1e-6), 1.0, (C, 192)))
    off, wmap = kernels.build_window_map(idx[:, 0] - 1, idx[:, 1], window, 2)
    out = kernels.element_transfer(chrom - 1, strand, ptr, bs, be, window, off, wmap, w64, yp, sd, yt, fl, dpr,
                                   L_elt=L, device=DEV)
    o = {k: v.cpu().numpy() for k, v in out.items()}
    for ci in range(C):
        want = oracle.gene_transfer(chrom, strand, np.array(ptr), np.array(bs), np.array(be), L, window, win_index,
                                    w64h, yp[ci], sd[ci], yt[ci], fl[ci], dpr[ci])
        np.testing.assert_allclose(o["MU"][ci], want["MU"], rtol=1e-12)
        np.testing.assert_allclose(o["SIGMA"][ci], want["SIGMA"], rtol=1e-12)
        assert np.array_equal(o["R_OBS"][ci], want["R_OBS"])
        assert np.array_equal(o["FLAG"][ci].astype(bool), want["FLAG"])
        for j, k in enumerate(("P_SILENT", "P_MIS", "P_NONS", "P_SPLICE")):
            np.testing.assert_allclose(o["P"][ci, :, j], want[k], rtol=1e-9)
    assert np.array_equal(o["R_SIZE"], want["R_SIZE"])
    assert np.array_equal(o["ELT_SIZE"], want["GENE_LENGTH"])
    assert np.array_equal(o["N_WIN"], want["N_WIN"])


def test_fused_gene_test_kernels_golden():
    """dig_gene_scale_sums + dig_gene_burden_test (13 NB tests + Fisher per gene in one launch) against the
    reference's transfer_tools functions run on the same table (tests/golden/genes.npz)."""
    import torch
    from digdriver_b200 import kernels
    z = golden("genes")
    genes = list(z["pre_genes"])
    t = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(DEV, dtype=dt)
    mu, sigma = t(z["pre_MU"]), t(z["pre_SIGMA"])
    P = t(np.stack([z["pre_Pi_SYN"], z["pre_Pi_MIS"], z["pre_Pi_NONS"], z["pre_Pi_SPL"]], axis=1))
    pi_indel = t(z["pre_Pi_INDEL"])
    obs = t(np.stack([z["out_OBS_" + c] for c in ("SYN", "MIS", "NONS", "SPL", "INDEL")], axis=1), torch.int64)
    nsamp = t(np.stack([z["out_N_SAMP_" + c] for c in ("SYN", "MIS", "NONS", "SPL", "TRUNC", "NONSYN", "INDEL")], axis=1),
              torch.int64)
    cgc = np.array([g in ("TP53", "KRAS", "PIK3CA", "BRAF", "PTEN") for g in genes])
    sums = kernels.gene_scale_sums(mu, sigma, P, pi_indel, obs, cgc, genes.index("TP53"))
    s = sums.cpu().numpy()
    keep = np.array([g != "TP53" for g in genes])
    np.testing.assert_allclose(s[0], (z["pre_MU"][keep] * z["pre_Pi_SYN"][keep]).sum(), rtol=1e-13)
    np.testing.assert_allclose(s[1], (z["pre_Pi_INDEL"] * z["pre_ALPHA_INDEL"] * z["pre_THETA_INDEL"])[~cgc].sum(), rtol=1e-13)
    assert s[2] == z["out_OBS_INDEL"][~cgc].sum()
    out = kernels.gene_burden_test(mu, sigma, P, pi_indel, obs, nsamp, sums, 0.0, scale_factor=float(z["cj"])).cpu().numpy()
    res = dict(zip(kernels.GENE_OUT_ROWS, out))
    for name, got in res.items():
        key = "out_" + name
        if key not in z.files:
            continue
        if name.startswith("PVAL_"):
            assert_pvals_close(got, z[key])
        else:
            np.testing.assert_allclose(got, z[key], rtol=1e-12, err_msg=name)
    np.testing.assert_allclose(res["ALPHA"], z["pre_ALPHA"], rtol=0)
    np.testing.assert_allclose(res["Pi_TRUNC"], z["pre_Pi_TRUNC"], rtol=0)
    np.testing.assert_allclose(res["Pi_NONSYN"], z["pre_Pi_NONSYN"], rtol=1e-15)
    # scale by expectation: cj = n_syn / sums[0]
    out2 = kernels.gene_burden_test(mu, sigma, P, pi_indel, obs, nsamp, sums, 1234.0).cpu().numpy()
    np.testing.assert_allclose(out2[22], z["pre_THETA"] * (1234.0 / s[0]), rtol=1e-14)
    # a gene without countable context has P = NaN; the reference's sums are pandas Series.sum(), which skips NaN
    # (transfer_tools.py:814, :699): such a row must not turn the cohort-wide scale factors into NaN
    import pandas as pd
    P2, pi2 = P.clone(), pi_indel.clone()
    P2[3, :] = float("nan")
    pi2[5] = float("nan")
    s2 = kernels.gene_scale_sums(mu, sigma, P2, pi2, obs, cgc, genes.index("TP53")).cpu().numpy()
    syn2 = z["pre_Pi_SYN"].copy(); syn2[3] = np.nan
    ind2 = z["pre_Pi_INDEL"].copy(); ind2[5] = np.nan
    np.testing.assert_allclose(s2[0], pd.Series((z["pre_MU"] * syn2)[keep]).sum(), rtol=1e-13)
    np.testing.assert_allclose(s2[1], pd.Series((ind2 * z["pre_ALPHA_INDEL"] * z["pre_THETA_INDEL"])[~cgc]).sum(), rtol=1e-13)
    assert np.isfinite(s2).all()


def test_secondary_gene_tests_golden():
    """f-4: dig_gene_dnds_sel / dig_selection_coefficient and the DataFrame-level functions of transfer_tools against
    the outputs of the unmodified reference (gene_expected_muts_dnds, gene_pvalue_burden_dnds, gene_pvalue_sel_nb,
    selection_coefficient).  Expectations rel. 1e-9 (measured: bit-identical), p-values |dlog10 p| <= 1e-6."""
    import pandas as pd
    from digdriver_b200 import kernels
    from digdriver_b200.driver_model import transfer_tools as tt
    dev = torch.device("cuda:0")
    z = golden("secondary")
    cls = kernels.DNDS_CLASSES
    pi6 = np.stack([z["in_Pi_" + c] for c in cls], axis=1)
    obs6 = np.stack([z["in_OBS_" + c] for c in cls], axis=1)
    out = kernels.gene_dnds_sel(z["in_ALPHA"], z["in_THETA"], pi6, obs6, dev).cpu().numpy()
    for name, row in zip(kernels.DNDS_OUT_ROWS, out):
        want = z["out_" + name]
        if name.startswith("PVAL"):
            assert_pvals_close(row, want)
        else:
            assert np.array_equal(np.isnan(row), np.isnan(want)), name
            m = ~np.isnan(want)
            assert np.allclose(row[m], want[m], rtol=1e-9, atol=0), name
    # DataFrame-level API, called in the reference's order
    df = pd.DataFrame({k[3:]: z[k] for k in z.files if k.startswith("in_")})
    df = tt.gene_expected_muts_dnds(df)
    df = tt.gene_pvalue_burden_dnds(df)
    df = tt.gene_pvalue_sel_nb(df)
    for c in ("SYN", "MIS", "TRUNC"):
        tt.selection_coefficient(df, c, pvalue=True)
    for k in z.files:
        if not k.startswith("out_"):
            continue
        got, want = df[k[4:]].values, z[k]
        if k[4:].startswith("PVAL"):
            assert_pvals_close(got, want)
        else:
            m = ~np.isnan(want)
            assert np.array_equal(np.isnan(got), np.isnan(want)) and np.allclose(got[m], want[m], rtol=1e-9, atol=0), k
