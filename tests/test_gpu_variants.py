"""GPU parity tests for the rest of the reference's API on the path: the other p-value conventions of nb_model.py, the
log-likelihood terms / row-level LLR tests / gamma-Poisson selection test / indel burden by transfer of
transfer_tools.py (goldens from the unmodified reference, tests/golden/make_golden_variants.py), and the
bedtools-intersect front ends of mutation_tools.py on the overlap-join kernel (checked against a brute-force
restatement of `bedtools intersect`: bedtools itself is not available, so that part is parity-unpinned).
Tolerances: p-values |dlog10 p| <= 1e-6, expectations rel. error <= 1e-9, joins exact."""
import gzip
import os

import numpy as np
import pandas as pd
import pytest

from conftest import golden, assert_pvals_close

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------ nb_model p-value family

@pytest.mark.parametrize("fn,key,with_mu", [
    ("nb_pvalue_greater", "v_greater", False), ("nb_pvalue_greater_midp_DEPRECATED", "v_greater_midp_deprecated", False),
    ("nb_pvalue_less", "v_less", False), ("nb_pvalue_less_midp", "v_less_midp", False),
    ("nb_pvalue_midp", "v_midp", False), ("nb_pvalue_midp", "v_midp_mu", True), ("nb_pvalue_exact", "v_exact_mu", True)])
def test_pvalue_conventions_golden(fn, key, with_mu):
    from digdriver_b200.sequence_model import nb_model
    z = golden("variants")
    f = getattr(nb_model, fn)
    got = f(z["v_k"], z["v_alpha"], z["v_p"], z["v_mu"]) if with_mu else f(z["v_k"], z["v_alpha"], z["v_p"])
    assert_pvals_close(got, z[key])
    # scalar call, as the reference's loops make it
    i = 17
    one = f(float(z["v_k"][i]), float(z["v_alpha"][i]), float(z["v_p"][i]), float(z["v_mu"][i])) if with_mu else \
        f(float(z["v_k"][i]), float(z["v_alpha"][i]), float(z["v_p"][i]))
    assert isinstance(one, float)
    assert_pvals_close([one], [z[key][i]])


def test_exact_variant_is_bit_identical_to_k8_scalar_kernel():
    from digdriver_b200 import kernels
    z = golden("variants")
    a = kernels.nb_pvalue_variant("exact", z["v_k"], z["v_alpha"], z["v_p"]).cpu().numpy()
    b = kernels.nb_pvalue_exact(z["v_k"], z["v_alpha"], z["v_p"]).cpu().numpy()
    assert np.array_equal(a, b, equal_nan=True)


def test_pvalue_variant_bad_inputs_and_empty():
    from digdriver_b200 import kernels
    k = np.array([np.nan, 1.0, -1.0, 2.0, 2.0])
    a = np.array([1.0, np.nan, 1.0, -1.0, 1.0])
    p = np.array([0.5, 0.5, 0.5, 0.5, 1.5])
    for mode in kernels.NB_MODES:
        assert np.all(np.isnan(kernels.nb_pvalue_variant(mode, k, a, p).cpu().numpy()))
        assert kernels.nb_pvalue_variant(mode, np.zeros(0), np.zeros(0), np.zeros(0)).numel() == 0


# ------------------------------------------------------------------ transfer_tools: log-likelihoods, LLR tests

def _secondary_df():
    s = golden("secondary")
    df = pd.DataFrame({c[3:]: s[c] for c in s.files if c.startswith("in_")})
    for c in ("T_SYN", "MRFOLD", "EXP_SYN"):
        df[c] = s["out_" + c]
    return df


def _close_ll(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    inf = np.isinf(want)
    assert np.array_equal(got[inf], want[inf])
    m = np.isfinite(want)
    np.testing.assert_allclose(got[m], want[m], rtol=1e-10, atol=1e-10)


def test_loglik_terms_golden():
    from digdriver_b200.driver_model import transfer_tools as tt
    z, df = golden("variants"), _secondary_df()
    _close_ll(tt._ll_nb(df.OBS_MIS.values, df.ALPHA.values, (df.THETA * df.Pi_MIS).values), z["ll_nb"])
    _close_ll(tt._ll_pois(df.OBS_MIS.values, (df.ALPHA * df.THETA * df.Pi_MIS).values), z["ll_pois"])
    _close_ll(tt._ll_pois(df.OBS_NONS.values, df.OBS_NONS.values), z["ll_pois_self"])
    _close_ll(tt._ll_gamma(df.T_SYN.values, df.ALPHA.values, (df.THETA * df.Pi_SYN * df.MRFOLD).values), z["ll_gamma"])
    assert isinstance(tt._ll_pois(3, 2.5), float)
    np.testing.assert_allclose(tt._ll_pois(3, 2.5), 3 * np.log(2.5) - 2.5 - np.log(6.0), rtol=1e-13)


def test_llr_rows_and_gamma_poisson_golden():
    from digdriver_b200.driver_model import transfer_tools as tt
    z, df = golden("variants"), _secondary_df()
    for i in (0, 11, 37, 222):                                 # row-level calls (k = 0 rows, a zero target size)
        assert_pvals_close(tt._llr_test_nb(df.iloc[i]), z["llr_nb_rows"][i])
        assert_pvals_close(tt._llr_test_gamma_poiss(df.iloc[i]), z["llr_pg_rows"][i])
    out = tt.gene_pvalue_sel_gamma(df.copy())
    for j, c in enumerate(("SYN", "MIS", "NONS", "NONSYN")):
        assert_pvals_close(out["PVAL_%s_SEL_PG" % c].values, z["PVAL_%s_SEL_PG" % c])
        assert_pvals_close(out["PVAL_%s_SEL_PG" % c].values, z["llr_pg_rows"][:, j])
    # host scalar helpers
    got_t = np.array([tt._mle_t(r.OBS_SYN, 1, r.ALPHA, r.THETA * r.Pi_SYN) for _, r in df.iterrows()])
    got_m = np.array([tt._mrfold_factor(r.T_SYN, r.EXP_SYN) for _, r in df.iterrows()])
    assert np.array_equal(got_t, z["mle_t"]) and np.array_equal(got_m, z["mrfold"])


def test_gene_pvalue_indel_by_transfer_golden(tmp_path, monkeypatch):
    from digdriver_b200.driver_model import transfer_tools as tt
    z = golden("variants")
    with gzip.open(tmp_path / "dndscv_gene_cds.bed.gz", "wt") as fh:
        for c, s, e, g in zip(z["ind_cds_chrom"], z["ind_cds_start"], z["ind_cds_end"], z["ind_cds_gene"]):
            fh.write("%s\t%d\t%d\t%s\n" % (c, s, e, g))
    (tmp_path / "genes_CGC_ALL.txt").write_text("\n".join(z["ind_cgc"]) + "\n")
    monkeypatch.setenv("DIG_DATA_DIR", str(tmp_path))
    df = pd.DataFrame({"ALPHA": z["ind_ALPHA"], "THETA": z["ind_THETA"], "R_SIZE": z["ind_R_SIZE"],
                       "OBS_INDEL": z["ind_OBS_INDEL"]}, index=[str(g) for g in z["ind_genes"]])
    out = tt.gene_pvalue_indel_by_transfer(df)
    for c in ("LENGTH", "Pi_INDEL", "THETA_INDEL", "EXP_INDEL"):
        np.testing.assert_allclose(out[c].values.astype(np.float64), z["ind_out_" + c], rtol=1e-9, equal_nan=True)
    assert_pvals_close(out.PVAL_INDEL_BURDEN.values, z["ind_out_PVAL_INDEL_BURDEN"])


def test_scale_factor_attrs_and_deprecated_burden(tmp_path):
    from digdriver_b200 import storage
    from digdriver_b200.driver_model import transfer_tools as tt
    st = storage.Store(str(tmp_path / "pre"), "w")
    st.set_attrs(N_MUT_CDS=200, N_SAMPLES=8)
    df = pd.DataFrame({"SAMPLE": ["a", "b", "a", "c"]})
    assert tt.scale_factor_by_cds(str(tmp_path / "pre"), df) == 4 / 200
    assert tt.scale_factor_by_samples(str(tmp_path / "pre"), df) == 3 / 8
    z = golden("nbtest")
    m = pd.DataFrame({"OBS_SNV": z["elt_OBS_SNV"], "ALPHA": z["elt_out_ALPHA"], "THETA": z["elt_out_THETA"],
                      "Pi_SUM": z["elt_Pi_SUM"]})
    assert_pvals_close(tt.element_pvalue_burden_nb_DEPRECATED(m).PVAL_SNV_BURDEN.values, z["elt_out_PVAL_SNV_BURDEN"])


# ------------------------------------------------------------------ overlap join and its front ends

def _brute_pairs(ac, as_, ae, bc, bs, be):
    """bedtools intersect -wa -wb restated: all (a, b) with the same chromosome label, a.start < b.end, b.start < a.end."""
    pairs = set()
    for i in range(len(ac)):
        hit = np.flatnonzero((bc == ac[i]) & (as_[i] < be) & (bs < ae[i]))
        pairs.update((i, int(j)) for j in hit)
    return pairs


def _synthetic_join(seed, n_mut=3000, n_blk=700):
    rng = np.random.default_rng(seed)
    bc = rng.choice(np.array(["1", "2", "10", "X", "chr3"]), n_blk)
    bs = rng.integers(0, 50_000, n_blk)
    be = bs + rng.choice([1, 5, 200, 3000, 40_000], n_blk)               # nested and long blocks exercise the walk
    ac = rng.choice(np.array(["1", "2", "10", "X", "7"]), n_mut)
    as_ = rng.integers(0, 52_000, n_mut)
    ae = as_ + rng.choice([1, 1, 1, 2, 30], n_mut)
    return ac, as_, ae, bc, bs, be


def test_overlap_pairs_match_brute_force():
    from digdriver_b200.data_tools import mutation_tools as mt
    for seed in (1, 2):
        ac, as_, ae, bc, bs, be = _synthetic_join(seed)
        im, ib = mt._overlap_join(ac, as_, ae, bc, bs, be)
        assert len(im) == len(set(zip(im.tolist(), ib.tolist())))
        assert set(zip(im.tolist(), ib.tolist())) == _brute_pairs(ac, as_, ae, bc, bs, be)
        assert np.all(np.diff(im) >= 0)                                  # grouped by mutation, input order
    # empty sides
    e = np.zeros(0, dtype=np.int64)
    im, ib = mt._overlap_join(np.zeros(0, dtype=str), e, e, bc, bs, be)
    assert len(im) == 0 and len(ib) == 0
    im, ib = mt._overlap_join(ac, as_, ae, np.zeros(0, dtype=str), e, e)
    assert len(im) == 0 and len(ib) == 0


def test_restrict_mutations_to_regions():
    from digdriver_b200.sequence_model import sequence_tools as st
    mut = pd.DataFrame({"CHROM": [1, 1, 1, 2, 2], "START": [5, 10, 19, 5, 30], "END": [6, 11, 25, 6, 31],
                        "REF": list("AAAAA"), "ALT": list("CCCCC")})
    got = st.restrict_mutations_to_regions(mut, np.array([[1, 10, 20], [2, 0, 10]]))
    assert got.START.tolist() == [10, 19, 5]


def _write_join_files(tmp_path, seed):
    rng = np.random.default_rng(seed)
    n_mut, n_elt = 1500, 120
    chrom = rng.integers(1, 4, n_mut)
    start = rng.integers(0, 30_000, n_mut)
    is_indel = rng.random(n_mut) < 0.1
    end = start + np.where(is_indel, rng.integers(1, 6, n_mut), 1)
    mut = pd.DataFrame({0: chrom.astype(str), 1: start, 2: end, 3: rng.choice(list("ACGT"), n_mut),
                        4: rng.choice(list("ACGT"), n_mut), 5: ["S%d" % s for s in rng.integers(0, 25, n_mut)],
                        6: ".", 7: np.where(is_indel, "INDEL", "Noncoding"), 8: "A>C", 9: "AAA"})
    mut = pd.concat([mut, mut.iloc[:40]])                                # duplicate rows
    f_mut = str(tmp_path / "mut.tsv")
    mut.to_csv(f_mut, sep="\t", header=False, index=False)
    rows = []
    for e in range(n_elt):
        c = int(rng.integers(1, 4))
        s0 = int(rng.integers(0, 28_000))
        nb = int(rng.integers(1, 4))
        sizes = rng.integers(20, 400, nb)
        gaps = rng.integers(0, 300, nb)
        starts = np.cumsum(np.concatenate([[0], (sizes + gaps)[:-1]]))
        rows.append((str(c), s0, s0 + int(starts[-1] + sizes[-1]), "E%d" % e, 0, rng.choice(["+", "-"]), s0, s0, 0, nb,
                     ",".join(map(str, sizes)) + ",", ",".join(map(str, starts)) + ","))
    f_bed = str(tmp_path / "elts.bed")
    pd.DataFrame(rows).to_csv(f_bed, sep="\t", header=False, index=False)
    return f_mut, f_bed, mut.reset_index(drop=True), rows


def _blocks(rows):
    out = []
    for r in rows:
        sizes = [int(x) for x in r[10].strip(",").split(",")]
        starts = [int(x) for x in r[11].strip(",").split(",")]
        for sz, st in zip(sizes, starts):
            out.append((r[0], r[1] + st, r[1] + st + sz, r[3], r[5]))
    return pd.DataFrame(out, columns=["CHROM", "START", "END", "ELT", "STRAND"])


def test_bedtools_front_ends(tmp_path):
    from digdriver_b200.data_tools import mutation_tools as mt
    f_mut, f_bed, mut, rows = _write_join_files(tmp_path, 5)
    blk = _blocks(rows)
    pairs = sorted(_brute_pairs(mut[0].values, mut[1].values, mut[2].values, blk.CHROM.values, blk.START.values,
                                blk.END.values))
    # mutations_by_element: one row per (mutation row, block)
    got = mt.mutations_by_element(f_mut, f_bed, bed12=True)
    want = sorted((int(mut[1][i]), str(mut[5][i]), blk.ELT[j]) for i, j in pairs)
    assert sorted(zip(got.START.astype(int), got.SAMPLE.astype(str), got.ELT)) == want
    assert list(got.columns) == ['CHROM', 'START', 'END', 'REF', 'ALT', 'SAMPLE', 'GENE', 'ANNOT', 'TYPE', 'CONTEXT', 'ELT']
    dd = mt.mutations_by_element(f_mut, f_bed, bed12=True, drop_duplicates=True)
    assert len(dd) == len(got.drop_duplicates(['CHROM', 'START', 'END', 'REF', 'ALT', 'SAMPLE', 'ELT']))
    # restrict_mutations_by_bed_efficient: -wa copies, then read_mutation_file's clean-up
    eff = mt.restrict_mutations_by_bed_efficient(f_mut, f_bed, bed12=True, drop_duplicates=True)
    hit_rows = mut.iloc[sorted({i for i, _ in pairs})]
    snv = hit_rows[hit_rows[7] != "INDEL"].drop_duplicates([0, 1, 2, 3, 4, 5])
    ind = hit_rows[hit_rows[7] == "INDEL"].drop_duplicates([0, 1, 2, 3, 4, 5]).drop_duplicates([0, 1, 2, 3, 4, 6])
    assert len(eff) == len(snv) + len(ind)
    assert set(zip(eff.CHROM.astype(str), eff.START, eff.SAMPLE)) == \
        set(zip(snv[0], snv[1], snv[5])) | set(zip(ind[0], ind[1], ind[5]))
    f_bed3 = str(tmp_path / "blocks.bed")                                  # the same blocks as a plain 3-column bed
    blk[["CHROM", "START", "END"]].to_csv(f_bed3, sep="\t", header=False, index=False)
    eff3 = mt.restrict_mutations_by_bed_efficient(f_mut, f_bed3, bed12=False, drop_duplicates=True)
    assert len(eff3) == len(eff) and set(zip(eff3.CHROM.astype(str), eff3.START, eff3.SAMPLE)) == \
        set(zip(eff.CHROM.astype(str), eff.START, eff.SAMPLE))
    nodrop = mt.restrict_mutations_by_bed_efficient(f_mut, f_bed, bed12=True, drop_duplicates=False, drop_sex=False)
    assert len(nodrop[nodrop.ANNOT != "INDEL"]) == sum(1 for i, _ in pairs if mut[7][i] != "INDEL")   # one copy per pair
    # restrict_mutations_by_bed: clipped coordinates, pybedtools column names, duplicates dropped
    df_bed = blk[["CHROM", "START", "END"]]
    cl = mt.restrict_mutations_by_bed(mut, df_bed, unique=False)
    assert list(cl.columns[:6]) == ['chrom', 'start', 'end', 'name', 'score', 'strand'] and len(cl) == len(pairs)
    want_cl = sorted((max(int(mut[1][i]), int(blk.START[j])), min(int(mut[2][i]), int(blk.END[j]))) for i, j in pairs)
    assert sorted(zip(cl.start.astype(int), cl.end.astype(int))) == want_cl
    assert len(mt.restrict_mutations_by_bed(mut, df_bed)) == len(cl.drop_duplicates())
    # tabulate_nonc_mutations_split: per element distinct samples and overlapping base pairs
    none, whole = mt.tabulate_nonc_mutations_split(f_bed, f_mut)
    assert none is None and list(whole.columns) == ['CHROM', 'ELT', 'STRAND', 'BLOCK_STARTS', 'BLOCK_ENDS',
                                                    'OBS_SAMPLES', 'OBS_MUT']
    dmut = mut.drop_duplicates([0, 1, 2, 3, 4, 5])
    dsnv, dind = dmut[dmut[7] != "INDEL"], dmut[dmut[7] == "INDEL"].drop_duplicates([0, 1, 2, 3, 4, 6])
    dm = pd.concat([dsnv, dind]).reset_index(drop=True)
    for r in whole.itertuples(index=False):
        b = blk[blk.ELT == r.ELT]
        samples, bp = set(), 0
        for bb in b.itertuples(index=False):
            h = dm[(dm[0] == bb.CHROM) & (dm[1] < bb.END) & (dm[2] > bb.START)]
            samples.update(h[5])
            bp += int((np.minimum(h[2], bb.END) - np.maximum(h[1], bb.START)).sum())
        assert (r.OBS_SAMPLES, r.OBS_MUT) == (len(samples), bp), r.ELT
        assert r.BLOCK_STARTS == sorted(set(b.START)) and r.BLOCK_ENDS == sorted(set(b.END))
    cols = pd.DataFrame({"Missense": [1]})
    mt._genic_fill_empty_cols(cols)
    assert set(cols.columns) == {'Essential_Splice', 'Missense', 'Nonsense', 'Stop_loss', 'Synonymous'}


def test_overlap_join_at_scale_counts_identity():
    """1 M mutations x 300 k blocks (nested, long and unit-length): the number of blocks a mutation overlaps equals
    #{blocks starting before its end} - #{blocks ending at or before its start}, two sorts on the host; the filled
    pairs all satisfy the overlap rule and agree with the counts."""
    from digdriver_b200 import kernels
    rng = np.random.default_rng(12)
    n_blk, n_mut = 300_000, 1_000_000
    chrom_b = rng.integers(0, 22, n_blk).astype(np.int64)
    bs = rng.integers(0, 100_000_000, n_blk)
    be = bs + rng.choice([1, 200, 2000, 50_000, 2_000_000], n_blk, p=[0.05, 0.4, 0.4, 0.14, 0.01])
    chrom_m = rng.integers(0, 23, n_mut).astype(np.int64)                  # chromosome 22 has no blocks
    ms = rng.integers(0, 100_000_000, n_mut)
    me = ms + rng.choice([1, 1, 1, 4], n_mut)
    kbs, kbe, kms, kme = (chrom_b << 32) | bs, (chrom_b << 32) | be, (chrom_m << 32) | ms, (chrom_m << 32) | me
    cnt = kernels.overlap_counts(kbs, kbe, kms, kme)
    want = np.searchsorted(np.sort(kbs), kme, side="left") - np.searchsorted(np.sort(kbe), kms, side="right")
    assert np.array_equal(cnt, want)
    im, ib = kernels.overlap_pairs(kbs, kbe, kms, kme)
    assert len(im) == int(want.sum()) and np.array_equal(np.bincount(im, minlength=n_mut), want)
    assert np.all((kms[im] < kbe[ib]) & (kbs[ib] < kme[im]))
    assert len(np.unique(im * n_blk + ib)) == len(im)


def test_region_counts_and_psum_match_k6_at_scale():
    """100 k elements on a 310 k-window map: dig_element_region_counts sums to K6's R_SIZE, and dig_element_psum on
    those counts reproduces K6's P bit for bit (same lane assignment and summation order)."""
    import torch
    from digdriver_b200 import kernels
    rng = np.random.default_rng(21)
    W, n_chrom, per = 10_000, 22, 14_000
    chrom = np.repeat(np.arange(n_chrom), per)
    start = np.tile(np.arange(per) * W, n_chrom)
    off, wmap = kernels.build_window_map(chrom, start, W, n_chrom)
    n_win = len(chrom)
    wc = rng.integers(0, 400, (n_win, 64)).astype(np.int32)
    wc[rng.random(n_win) < 0.01] = 0                                        # all-N windows
    E = 100_000
    nb = rng.integers(1, 4, E)
    ptr = np.concatenate([[0], np.cumsum(nb)])
    ec = rng.integers(0, n_chrom, E).astype(np.int32)
    first = rng.integers(0, (per - 5) * W, E)
    owner = np.repeat(np.arange(E), nb)
    step = rng.integers(300, 9000, len(owner))
    rel = np.cumsum(step) - step
    rel -= rel[ptr[:-1]][owner]
    bs = first[owner] + rel
    be = bs + rng.integers(200, 2000, len(bs))
    es = rng.choice([-1, 1], E).astype(np.int8)
    L = rng.integers(0, 30, (E, 192)).astype(np.float64)
    d_pr = rng.lognormal(np.log(1e-6), 1.0, 192)
    rp = [rng.gamma(2.0, 10.0, n_win), rng.uniform(0.5, 5, n_win), rng.poisson(20, n_win).astype(float), rng.random(n_win) < 0.1]
    pre = kernels.element_transfer(ec, es, ptr, bs, be, W, off, wmap, wc, rp[0], rp[1], rp[2], rp[3], d_pr,
                                   L_elt=L.reshape(E, 192, 1))
    rc, nw = kernels.element_region_counts(ec, es, ptr, bs, be, W, off, wmap, wc)
    assert torch.equal(rc.sum(dim=1), pre["R_SIZE"]) and torch.equal(nw, pre["N_WIN"])
    p = kernels.element_psum(L, torch.repeat_interleave(rc, 3, dim=1), d_pr)
    got, want = p.cpu().numpy(), pre["P"][0][:, 0].cpu().numpy()
    assert np.array_equal(got, want, equal_nan=True)
    # strand: a minus-strand element's counts are the plus-strand counts re-indexed by the reverse complement
    rc_plus, _ = kernels.element_region_counts(ec, np.ones(E, dtype=np.int8), ptr, bs, be, W, off, wmap, wc)
    comp = np.array([((3 - (c & 3)) << 4) | ((3 - ((c >> 2) & 3)) << 2) | (3 - ((c >> 4) & 3)) for c in range(64)])
    a, b = rc.cpu().numpy(), rc_plus.cpu().numpy()
    minus = es < 0
    assert np.array_equal(a[~minus], b[~minus]) and np.array_equal(a[minus], b[minus][:, comp])


def test_new_entry_points_accept_empty_inputs():
    from digdriver_b200 import kernels
    z = np.zeros(0)
    zi = np.zeros(0, dtype=np.int64)
    assert kernels.loglik("pois", z, z).numel() == 0 and kernels.loglik("gamma", z, z, z).numel() == 0
    assert kernels.gene_llr_test("nb", z, z, np.zeros((0, 3)), np.zeros((0, 3)), z).shape == (4, 0)
    assert kernels.gene_llr_test("gamma_poisson", z, z, np.zeros((0, 3)), np.zeros((0, 3)), z, z).shape == (4, 0)
    assert kernels.element_psum(np.zeros((0, 192)), np.zeros((0, 192), dtype=np.int64), np.ones(192)).numel() == 0
    off, wmap = kernels.build_window_map(np.array([0]), np.array([0]), 1000, 1)
    rc, nw = kernels.element_region_counts(np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int8), np.zeros(1, dtype=np.int64),
                                           zi, zi, 1000, off, wmap, np.zeros((1, 64), dtype=np.int32))
    assert rc.shape == (0, 64) and nw.numel() == 0
    assert kernels.overlap_counts(zi, zi, zi, zi).size == 0
    with pytest.raises(Exception):
        kernels.gene_llr_test("gamma_poisson", np.ones(2), np.ones(2), np.ones((2, 3)), np.ones((2, 3)), np.ones(2))   # no T_SYN
