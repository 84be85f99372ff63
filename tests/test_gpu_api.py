"""End-to-end GPU tests of the reference-facing API: the DataFrame-level functions of sequence_model /
driver_model / data_tools and the three CLI shims, run on files, against the golden vectors and against an
independent composition of the CPU oracle on the same inputs."""
import os

import numpy as np
import pandas as pd
import pytest

from conftest import golden, assert_pvals_close, GOLDEN

pytestmark = pytest.mark.gpu

W = 1000


@pytest.fixture(scope="module", autouse=True)
def _data_dir():
    os.environ["DIG_DATA_DIR"] = os.path.join(GOLDEN, "data")
    yield


@pytest.fixture(scope="module")
def gold_genome():
    from digdriver_b200.genome import Genome
    z = golden("scan")
    return Genome(["chr1", "chr2"], [z["seq_chr1"], z["seq_chr2"]])


# ------------------------------------------------------------------ DataFrame-level API vs golden

def test_count_contexts_by_regions_golden(gold_genome):
    from digdriver_b200.sequence_model import sequence_tools as st
    z = golden("scan")
    for (u, d) in [(1, 1), (2, 2)]:
        w = z["windows"][z["rows_%d_%d" % (u, d)]]
        df = st.count_contexts_by_regions(gold_genome, ["chr%d" % c for c in w[:, 0]], w[:, 1], w[:, 2], n_up=u, n_down=d)
        assert list(df.columns) == list(z["columns_%d_%d" % (u, d)])
        assert list(df.index) == list(z["index_%d_%d" % (u, d)])
        assert np.array_equal(df.values, z["counts_%d_%d" % (u, d)])
    df_bed = pd.DataFrame(z["windows"][:60], columns=[0, 1, 2])
    df = st.count_contexts_in_bed(gold_genome, df_bed, n_up=1, n_down=1, N_proc=2, N_chunk=4)
    assert np.array_equal(df.values, z["pool_counts_1_1"]) and list(df.index) == list(z["pool_index"])
    with pytest.raises(ValueError):
        st.count_contexts_by_regions(gold_genome, ["chr1"], [1], [50], n_up=2, n_down=2)


def test_count_sequence_context_and_collapse(gold_genome, oracle):
    from digdriver_b200.sequence_model import sequence_tools as st
    seq = gold_genome.fetch("chr1", 100, 2100).upper()
    got = st.count_sequence_context(seq, n_up=1, n_down=1)
    want = oracle.py_count_sequence_context(seq, 1, 1)
    assert got == want and list(got.keys()) == list(want.keys())
    col = st.count_sequence_context(seq, n_up=1, n_down=1, collapse=True)
    assert len(col) == 32 and sum(col.values()) == sum(want.values())
    assert col["ACA"] == want["ACA"] + want["TGT"]


def test_nonc_elt_context_count_golden(gold_genome):
    from digdriver_b200.sequence_model import sequence_tools as st
    z = golden("scan")
    regions = [(int(c), int(s), int(e), "-" if t < 0 else "+") for c, s, e, t in
               zip(z["blk_chrom"], z["blk_start"], z["blk_end"], z["blk_strand"])]
    df = st.nonc_elt_context_count(regions, st.mk_trans_idx(), gold_genome)
    assert list(df.columns) == list(z["blk_columns"])
    assert np.array_equal(df.values, z["blk_L192"])


def test_mutation_contexts_by_chrom_golden(gold_genome):
    from digdriver_b200.sequence_model import sequence_tools as st
    z = golden("mutctx")
    df = pd.DataFrame({c: z["in_" + c] for c in ("CHROM", "START", "END", "REF", "ALT", "SAMPLE", "GENE", "ANNOT")})
    df["ROW"] = np.arange(len(df))
    for (u, d) in [(1, 1), (2, 2)]:
        outs = [st.mutation_contexts_by_chrom(gold_genome, g, n_up=u, n_down=d) for _, g in df.groupby("CHROM")]
        o = pd.concat(outs)
        assert np.array_equal(o.ROW.values, z["kept_rows_%d_%d" % (u, d)])
        assert list(o.CONTEXT) == list(z["context_%d_%d" % (u, d)])
        assert list(o.MUT_TYPE) == list(z["mut_type_%d_%d" % (u, d)])


def test_index_tables_match_reference():
    from digdriver_b200.sequence_model import sequence_tools as st
    z = golden("tables")
    assert st.mk_trans_idx() == list(z["trans_idx"])
    df = st.mk_mutation_context(return_df=True)
    assert list(df.MUT_TYPE) == list(z["mutctx_mut"]) and list(df.CONTEXT) == list(z["mutctx_ctx"])


def test_transfer_tools_element_functions_golden():
    from digdriver_b200.driver_model import transfer_tools as tt
    from digdriver_b200.sequence_model import nb_model
    z = golden("nbtest")
    n = len(z["elt_MU"])
    alpha, theta = nb_model.normal_params_to_gamma(z["elt_MU"], z["elt_SIGMA"])
    idx = ["E%d" % i for i in range(n)]
    df_pre = pd.DataFrame({"ELT_SIZE": 100, "FLAG": False, "R_SIZE": 10000, "R_OBS": 5, "R_INDEL": 5,
                           "MU": z["elt_MU"], "SIGMA": z["elt_SIGMA"], "ALPHA": alpha, "THETA": theta,
                           "MU_INDEL": z["elt_MU"], "SIGMA_INDEL": z["elt_SIGMA"], "ALPHA_INDEL": alpha,
                           "THETA_INDEL": theta, "Pi_SUM": z["elt_Pi_SUM"], "Pi_INDEL": z["elt_Pi_INDEL"]}, index=idx)
    df_tab = pd.DataFrame({"OBS_SAMPLES": z["elt_OBS_SAMPLES"], "OBS_SNV": z["elt_OBS_SNV"],
                           "OBS_INDEL": z["elt_OBS_INDEL"]}, index=idx)[: n - 100]
    dfm = tt.finish_element_model(df_tab, df_pre, float(z["elt_cj"]), float(z["elt_cj_indel"]))
    assert np.array_equal(dfm.EXP_SNV.values, z["elt_out_EXP_SNV"])
    np.testing.assert_allclose(dfm.EXP_INDEL.values, z["elt_out_EXP_INDEL"], rtol=1e-15)
    for c in ("PVAL_SNV_BURDEN", "PVAL_SAMPLE_BURDEN", "PVAL_INDEL_BURDEN", "PVAL_MUT_BURDEN"):
        assert_pvals_close(dfm[c].values, z["elt_out_" + c])
    assert np.all(dfm.OBS_SNV.values[-100:] == 0)


def test_transfer_tools_gene_functions_golden():
    from digdriver_b200.driver_model import transfer_tools as tt
    from digdriver_b200.data_tools import mutation_tools as mt
    z = golden("genes")
    gm = pd.DataFrame({c: z["in_" + c] for c in ("CHROM", "START", "END", "REF", "ALT", "SAMPLE", "GENE", "ANNOT")})
    pre = pd.DataFrame({c[4:]: z[c] for c in z.files if c.startswith("pre_") and c != "pre_genes"}, index=z["pre_genes"])
    cnt = mt.mutations_per_gene(gm)
    dfg = tt.transfer_gene_model(gm, cnt, pre, float(z["cj"]))
    dfg = tt.gene_expected_muts_nb(dfg)
    dfg = tt.gene_pvalue_burden_nb(dfg)
    dfg = tt.gene_pvalue_burden_nb_by_sample(dfg)
    dfg = tt.gene_pvalue_indel(dfg)
    dfg["PVAL_MUT_BURDEN"] = tt.fisher_combine(dfg.PVAL_TRUNC_BURDEN.values, dfg.PVAL_INDEL_BURDEN.values)
    for c in z.files:
        if not c.startswith("out_"):
            continue
        col = c[4:]
        if col.startswith("PVAL_"):
            assert_pvals_close(dfg[col].values, z[c])
        elif col.startswith(("OBS_", "N_SAMP_")):
            assert np.array_equal(dfg[col].values.astype(np.float64), z[c]), col
        else:
            np.testing.assert_allclose(dfg[col].values, z[c], rtol=1e-12, err_msg=col)


# ------------------------------------------------------------------ the CLI flow on files vs the oracle

def _write_fasta(path, seqs):
    with open(path, "w") as f:
        for name, s in seqs.items():
            f.write(">%s some description\n" % name)
            txt = s.tobytes().decode()
            for i in range(0, len(txt), 60):
                f.write(txt[i:i + 60] + "\n")


@pytest.fixture(scope="module")
def workdir(tmp_path_factory, oracle):
    """A small two-chromosome project laid out as the files the reference's CLI consumes."""
    d = tmp_path_factory.mktemp("proj")
    rng = np.random.default_rng(2026)
    lens = {"chr1": 60_000, "chr2": 45_500}
    seqs = {}
    for name, L in lens.items():
        s = rng.choice(np.frombuffer(b"ACGTacgt", dtype=np.uint8), size=L)
        a = int(rng.integers(1000, L - 3000))
        s[a:a + 700] = ord("N")
        seqs[name] = s
    _write_fasta(d / "genome.fa", seqs)
    wins = []
    for c, (name, L) in enumerate(lens.items(), start=1):
        i = 0
        while i + W < L:
            wins.append((c, i, i + W))
            i += W
    wins = np.array(wins)
    pd.DataFrame(wins).to_csv(d / "windows.bed", sep="\t", header=False, index=False)
    nW = len(wins)
    rp = pd.DataFrame({"CHROM": wins[:, 0], "START": wins[:, 1], "END": wins[:, 2],
                       "Y_TRUE": rng.poisson(20, nW), "Y_PRED": rng.gamma(2.0, 10.0, nW),
                       "STD": rng.uniform(0.5, 5.0, nW), "FLAG": rng.random(nW) < 0.1},
                      index=["chr%d:%d-%d" % tuple(r) for r in wins])
    # genes == elements (bed12), both strands, 1-4 blocks
    rows, genes = [], []
    for gi in range(40):
        c = int(rng.integers(1, 3))
        Lc = (list(lens.values())[c - 1] // W - 1) * W
        nb = int(rng.integers(1, 5))
        st = np.sort(rng.choice(np.arange(1500, Lc - 2500, 50), nb, replace=False))
        sz = rng.integers(20, 45, nb)
        name = "TP53" if gi == 3 else "GENE%02d" % gi
        strand = "-" if rng.random() < 0.5 else "+"
        rows.append((c, st[0], st[-1] + sz[-1], name, 0, strand, st[0], st[-1] + sz[-1], 0, nb,
                     ",".join(map(str, sz)) + ",", ",".join(map(str, st - st[0])) + ","))
        genes.append((name, c, strand, st, st + sz))
    pd.DataFrame(rows).to_csv(d / "elements.bed", sep="\t", header=False, index=False)
    # raw 8-column mutation file: half of the SNVs inside gene blocks
    up = {k: np.char.upper(v.tobytes().decode()).item() for k, v in seqs.items()}
    muts = []
    annots = ["Synonymous", "Missense", "Nonsense", "Essential_Splice"]
    for i in range(4000):
        if i % 2 == 0:
            name, c, strand, bs, be = genes[int(rng.integers(0, len(genes)))]
            b = int(rng.integers(0, len(bs)))
            p = int(rng.integers(bs[b], be[b]))
            gene, annot = name, annots[int(rng.integers(0, 4))]
        else:
            c = int(rng.integers(1, 3))
            p = int(rng.integers(5, list(lens.values())[c - 1] - 5))
            gene, annot = ".", "Noncoding"
        ref = up["chr%d" % c][p]
        if ref == "N":
            continue
        alt = "ACGT"[("ACGT".index(ref) + int(rng.integers(1, 4))) % 4]
        if rng.random() < 0.06:                                  # indel rows
            muts.append((c, p, p + int(rng.integers(2, 9)), ref + "AC", ref, "S%02d" % (rng.zipf(1.4) % 25), gene, "INDEL"))
        else:
            muts.append((c, p, p + 1, ref, alt, "S%02d" % (rng.zipf(1.4) % 25), gene, annot))
    muts += muts[:150]                                           # duplicated rows
    dfm = pd.DataFrame(muts).sort_values([0, 1], kind="stable")
    dfm.to_csv(d / "raw.tsv", sep="\t", header=False, index=False)
    # synthetic f_genic store
    from digdriver_b200 import storage
    gs = storage.Store(str(d / "genic"), "w")
    gs.write_table("genes", pd.DataFrame({"GENE": [g[0] for g in genes], "CHROM": [g[1] for g in genes],
                                          "STRAND": [g[2] for g in genes]}))
    ptr = np.concatenate([[0], np.cumsum([len(g[3]) for g in genes])])
    gs.write_array("cds_ptr", ptr)
    gs.write_array("cds_start", np.concatenate([g[3] for g in genes]))
    gs.write_array("cds_end", np.concatenate([g[4] - 1 for g in genes]))          # inclusive CDS ends
    Ldata = rng.integers(0, 30, (len(genes), 192, 4)).astype(np.float64)
    gs.write_array("L_data", Ldata)
    storage.Store(str(d / "pretrained"), "w").write_table("region_params", rp)
    return dict(dir=d, seqs=seqs, lens=lens, wins=wins, rp=rp, genes=genes, Ldata=Ldata)


def _cli(mod, text):
    import importlib.util
    import sys
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", mod + ".py")
    spec = importlib.util.spec_from_file_location("cli_" + mod, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    args = m.parse_args(text)
    args.func(args)


def test_cli_flow_matches_oracle(workdir, oracle):
    from digdriver_b200 import storage
    d = workdir["dir"]
    p = lambda x: str(d / x)
    # ---- 1. genome scan
    _cli("DigPreprocess", "countGenomeContext %s %s --bed %s --up 1 --down 1" % (p("genome.fa"), p("counts"), p("windows.bed")))
    cs = storage.Store(p("counts"), "r")
    df_counts = cs.read_table("all_window_genome_counts")
    off2 = 60_032
    seq = np.full(off2 + 45_500, ord("N"), dtype=np.uint8)
    seq[:60_000] = workdir["seqs"]["chr1"]
    seq[off2:] = workdir["seqs"]["chr2"]
    off, ln = np.array([0, off2]), np.array([60_000, 45_500])
    wins = workdir["wins"]
    want_counts, _ = oracle.count_regions(seq, off, ln, wins[:, 0] - 1, wins[:, 1], wins[:, 2], 1, 1)
    assert np.array_equal(df_counts.values, want_counts)
    assert list(df_counts.columns) == oracle.context_names(1, 1)
    assert np.array_equal(cs.read_table("genome_counts").values, want_counts.sum(axis=0))
    assert np.array_equal(cs.read_array("idx"), wins) and cs.get_attrs()["n_up"] == 1
    # ---- 2. mutation contexts
    _cli("DigPreprocess", "addMutationContext %s %s %s --up 1 --down 1" % (p("raw.tsv"), p("genome.fa"), p("annot.tsv")))
    ann = pd.read_table(p("annot.tsv"), header=None)
    assert ann.shape[1] == 10
    snv = ann[ann[7] != "INDEL"]
    names = np.array(oracle.context_names(1, 1))
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    ctx = oracle.mutation_contexts(seq, off, ln, snv[0].values - 1, snv[1].values,
                                   np.array([code.get(r, 255) for r in snv[3]], dtype=np.uint8), 1, 1)
    assert np.all(ctx >= 0) and list(names[ctx]) == list(snv[9])
    assert list(snv[8]) == ["%s>%s" % (r, a) for r, a in zip(snv[3], snv[4])]
    assert (ann[7] == "INDEL").sum() > 0 and set(ann[ann[7] == "INDEL"][9]) == {"."}
    # ---- 3. sequence model
    _cli("DigPretrain", "sequenceModel %s %s %s" % (p("annot.tsv"), p("counts"), p("pretrained")))
    pre = storage.Store(p("pretrained"), "r")
    m192 = pre.read_table("sequence_model_192")
    dd = ann.drop_duplicates([0, 1, 2, 3, 4, 5])
    dd = dd[dd[7] != "INDEL"]
    inwin = np.array([(int(c), int(s) // W * W) in {(int(a), int(b)) for a, b, _ in wins} for c, s in zip(dd[0], dd[1])])
    dd = dd[inwin].drop_duplicates()
    S = want_counts.sum(axis=0)
    cpos = {k: i for i, k in enumerate(names)}
    for (mt_, cx, cnt, fr) in zip(m192.MUT_TYPE, m192.CONTEXT, m192.COUNT, m192.FREQ):
        c = int(((dd[8] == mt_) & (dd[9] == cx)).sum())
        assert c == cnt
        assert fr == c / S[cpos[cx]]
    d_pr = oracle.model192_to_dpr(m192.FREQ.values)
    # ---- 4. gene model
    _cli("DigPretrain", "genicModel %s %s --fasta %s" % (p("pretrained"), p("genic"), p("genome.fa")))
    gm = storage.Store(p("pretrained"), "r").read_table("genic_model")
    genes = workdir["genes"]
    rp = workdir["rp"]
    win_index = {(int(c), int(s)): i for i, (c, s, e) in enumerate(wins)}
    g_chrom = np.array([g[1] for g in genes])
    g_strand = np.array([-1 if g[2] == "-" else 1 for g in genes], dtype=np.int8)
    ptr = np.concatenate([[0], np.cumsum([len(g[3]) for g in genes])])
    bs = np.concatenate([g[3] for g in genes])
    be_incl = np.concatenate([g[4] - 1 for g in genes])
    wantg = oracle.gene_transfer(g_chrom, g_strand, ptr, bs, be_incl, workdir["Ldata"], W, win_index, want_counts,
                                 rp.Y_PRED.values, rp.STD.values, rp.Y_TRUE.values.astype(float), rp.FLAG.values, d_pr)
    for col, key in (("MU", "MU"), ("SIGMA", "SIGMA"), ("P_MIS", "P_MIS"), ("P_NONS", "P_NONS"), ("P_SILENT", "P_SILENT"),
                     ("P_SPLICE", "P_SPLICE"), ("P_TRUNC", "P_TRUNC"), ("P_INDEL", "P_INDEL")):
        np.testing.assert_allclose(gm[col].values, wantg[key], rtol=1e-9, err_msg=col)
    assert np.array_equal(gm.R_SIZE.values, wantg["R_SIZE"]) and np.array_equal(gm.GENE_LENGTH.values, wantg["GENE_LENGTH"])
    # ---- 5. element model
    _cli("DigPreprocess", "initialize_f_data %s %s" % (p("eltdata"), p("counts")))
    _cli("DigPreprocess", "preprocess_element_model %s %s %s K1 --f-bed %s --window %d" %
         (p("eltdata"), p("pretrained"), p("genome.fa"), p("elements.bed"), W))
    _cli("DigPretrain", "elementModel %s %s K1" % (p("pretrained"), p("eltdata")))
    em = storage.Store(p("pretrained"), "r").read_table("K1")
    be_excl = np.concatenate([g[4] for g in genes])
    blk_elt = np.repeat(np.arange(len(genes)), np.diff(ptr))
    # the reference keeps only the FIRST row of duplicated 'chr:start-end' block keys (sequence_tools.py:525),
    # so a block shared by two elements is counted with the strand of its first occurrence
    blk_strand = g_strand[blk_elt].copy()
    first = {}
    for i, k in enumerate(zip(g_chrom[blk_elt], bs, be_excl)):
        blk_strand[i] = blk_strand[first.setdefault(k, i)]
    c64, _ = oracle.count_regions(seq, off, ln, g_chrom[blk_elt] - 1, bs, be_excl, 1, 1, strand=blk_strand)
    L = np.zeros((len(genes), 192))
    np.add.at(L, blk_elt, np.repeat(c64, 3, axis=1))
    wante = oracle.element_transfer(g_chrom, g_strand, ptr, bs, be_excl, L, W, win_index, want_counts, rp.Y_PRED.values,
                                    rp.STD.values, rp.Y_TRUE.values.astype(float), rp.FLAG.values, d_pr)
    assert list(em.ELT) == [g[0] for g in genes]
    for col in ("MU", "SIGMA", "P_SUM", "P_INDEL"):
        np.testing.assert_allclose(em[col].values, wante[col], rtol=1e-9, err_msg=col)
    for col in ("R_SIZE", "ELT_SIZE"):
        assert np.array_equal(em[col].values, wante[col]), col
    assert np.array_equal(em.R_OBS.values, wante["R_OBS"]) and np.array_equal(em.FLAG.values.astype(bool), wante["FLAG"])
    # ---- 6. element driver (default: scale by expected synonymous mutations) vs the oracle composition
    _cli("DigDriver", "elementDriver %s %s K1 --f-bed %s --outpfx elt --outdir %s" %
         (p("annot.tsv"), p("pretrained"), p("elements.bed"), p("out")))
    res = pd.read_table(p("out/elt.results.txt"), index_col=0)
    raw = pd.read_table(p("annot.tsv"), header=None)
    raw.columns = ["CHROM", "START", "END", "REF", "ALT", "SAMPLE", "GENE", "ANNOT", "MUT_TYPE", "CONTEXT"]
    blocks = pd.DataFrame({"CHROM": g_chrom[blk_elt], "START": bs, "END": be_excl,
                           "ELT": np.array([g[0] for g in genes])[blk_elt]})
    tab, black = oracle.tabulate_mutations_in_element(raw, blocks, max_muts_per_sample=3e9)
    cds = raw[raw.GENE != "."]
    cds = pd.concat([cds[cds.ANNOT != "INDEL"], cds[cds.ANNOT == "INDEL"].drop_duplicates(
        ["CHROM", "START", "END", "REF", "ALT", "GENE"])])           # read_mutation_file's unique_indels
    n_syn = len(cds[(cds.ANNOT == "Synonymous") & (cds.GENE != "TP53")].drop_duplicates())
    keep = np.array([g[0] != "TP53" for g in genes])
    cj = n_syn / (wantg["MU"][keep] * wantg["P_SILENT"][keep]).sum()
    a_g, t_g = oracle.normal_params_to_gamma(wantg["MU"], wantg["SIGMA"])
    null = np.array([g[0] not in ("TP53", "KRAS", "PIK3CA", "BRAF", "PTEN") for g in genes])
    cj_indel = (cds.ANNOT == "INDEL").sum() / (wantg["P_INDEL"][null] * a_g[null] * t_g[null]).sum()
    a_e, t_e = oracle.normal_params_to_gamma(wante["MU"], wante["SIGMA"])
    tab = tab.reindex([g[0] for g in genes]).fillna(0)
    exp, pv = oracle.burden_test(tab.OBS_SNV.values, a_e, t_e * cj, wante["P_SUM"])
    _, pvs = oracle.burden_test(tab.OBS_SAMPLES.values, a_e, t_e * cj, wante["P_SUM"])
    exp_i, pvi = oracle.burden_test(tab.OBS_INDEL.values, a_e, t_e * cj_indel, wante["P_INDEL"])
    res = res.loc[[g[0] for g in genes]]
    assert np.array_equal(res.OBS_SNV.values, tab.OBS_SNV.values) and np.array_equal(res.OBS_INDEL.values, tab.OBS_INDEL.values)
    assert np.array_equal(res.OBS_SAMPLES.values, tab.OBS_SAMPLES.values)
    np.testing.assert_allclose(res.EXP_SNV.values, exp, rtol=1e-9)
    np.testing.assert_allclose(res.EXP_INDEL.values, exp_i, rtol=1e-9)
    assert_pvals_close(res.PVAL_SNV_BURDEN.values, pv)
    assert_pvals_close(res.PVAL_SAMPLE_BURDEN.values, pvs)
    assert_pvals_close(res.PVAL_INDEL_BURDEN.values, pvi)
    assert_pvals_close(res.PVAL_MUT_BURDEN.values, oracle.fisher2(pv, pvi))
    # ---- 7. quickDriver == elementDriver except for the reference's double application of cj_indel
    _cli("DigDriver", "quickDriver %s %s %s --f_elts_bed %s --outpfx quick --outdir %s" %
         (p("annot.tsv"), p("pretrained"), p("genome.fa"), p("elements.bed"), p("out")))
    q = pd.read_table(p("out/quick.results.txt"), index_col=0).loc[[g[0] for g in genes]]
    np.testing.assert_allclose(q.EXP_SNV.values, exp, rtol=1e-9)
    assert_pvals_close(q.PVAL_SNV_BURDEN.values, pv)
    _, pvi2 = oracle.burden_test(tab.OBS_INDEL.values, a_e, t_e * cj_indel * cj_indel, wante["P_INDEL"])
    assert_pvals_close(q.PVAL_INDEL_BURDEN.values, pvi2)
    # ---- 8. gene driver
    _cli("DigDriver", "geneDriver %s %s --outpfx gene --outdir %s" % (p("annot.tsv"), p("pretrained"), p("out")))
    gres = pd.read_table(p("out/gene.results.txt"), index_col=0).loc[[g[0] for g in genes]]
    obs = oracle.gene_observed_counts(cds).reindex([g[0] for g in genes]).fillna(0)
    cjg = len(cds[(cds.GENE != "TP53") & (cds.ANNOT == "Synonymous")]) / (wantg["MU"][keep] * wantg["P_SILENT"][keep]).sum()
    pis = {"SYN": wantg["P_SILENT"], "MIS": wantg["P_MIS"], "NONS": wantg["P_NONS"], "SPL": wantg["P_SPLICE"]}
    pis["TRUNC"] = pis["NONS"] + pis["SPL"]
    pis["NONSYN"] = pis["MIS"] + pis["TRUNC"]
    ks = {c: obs["OBS_" + c].values.astype(float) for c in ("SYN", "MIS", "NONS", "SPL")}
    ks["TRUNC"] = ks["NONS"] + ks["SPL"]
    ks["NONSYN"] = ks["MIS"] + ks["TRUNC"]
    for c in ("SYN", "MIS", "NONS", "SPL", "TRUNC", "NONSYN"):
        e_, p_ = oracle.burden_test(ks[c], a_g, t_g * cjg, pis[c])
        assert np.array_equal(gres["OBS_" + c].values, ks[c]), c
        np.testing.assert_allclose(gres["EXP_" + c].values, e_, rtol=1e-9)
        assert_pvals_close(gres["PVAL_%s_BURDEN" % c].values, p_)
        _, ps_ = oracle.burden_test(obs["N_SAMP_" + c].values.astype(float), a_g, t_g * cjg, pis[c])
        assert_pvals_close(gres["PVAL_%s_BURDEN_SAMPLE" % c].values, ps_)
    t_ind = obs.OBS_INDEL.values[null].sum() / (wantg["P_INDEL"][null] * a_g[null] * t_g[null]).sum()
    _, pgi = oracle.burden_test(obs.OBS_INDEL.values.astype(float), a_g, t_g * t_ind, wantg["P_INDEL"])
    assert_pvals_close(gres.PVAL_INDEL_BURDEN.values, pgi)


def test_target_driver_cli(workdir, oracle, tmp_path):
    """DigDriver.py targetDriver (run_target_model, transfer_tools.py:876-967) on the project of the CLI flow test:
    panel restriction, panel scale factor N_MUT / N_MUT_<panel> from the archive attributes, burden tests."""
    from digdriver_b200 import storage
    d = workdir["dir"]
    p = lambda x: str(d / x)
    if not os.path.exists(p("annot.tsv")) or not storage.Store(p("pretrained"), "r").has("genic_model"):
        test_cli_flow_matches_oracle(workdir, oracle)
    genes = [g[0] for g in workdir["genes"]]
    panel = genes[::2]
    (tmp_path / "genes_MSK_341.txt").write_text("\n".join(panel) + "\n")
    storage.Store(p("pretrained"), "a").set_attrs(N_MUT_MSK_341=1234, N_SAMPLE_MSK_341=77, N_MUT_SAMPLE_MSK_341=999)
    old = os.environ.get("DIG_DATA_DIR")
    os.environ["DIG_DATA_DIR"] = str(tmp_path)
    try:
        _cli("DigDriver", "targetDriver %s %s --panel MSK_341 --outpfx targ --outdir %s" %
             (p("annot.tsv"), p("pretrained"), p("out")))
    finally:
        os.environ["DIG_DATA_DIR"] = old
    res = pd.read_table(p("out/targ.results.txt"), index_col=0)
    assert sorted(res.index) == sorted(panel)
    res = res.loc[panel]
    raw = pd.read_table(p("annot.tsv"), header=None)
    raw.columns = ["CHROM", "START", "END", "REF", "ALT", "SAMPLE", "GENE", "ANNOT", "MUT_TYPE", "CONTEXT"]
    cds = raw[raw.GENE != "."]
    cds = pd.concat([cds[cds.ANNOT != "INDEL"], cds[cds.ANNOT == "INDEL"].drop_duplicates(
        ["CHROM", "START", "END", "REF", "ALT", "GENE"])])
    cds = cds[cds.GENE.isin(panel)]
    dd = raw.drop_duplicates(["CHROM", "START", "END", "REF", "ALT", "SAMPLE"])
    dd = pd.concat([dd[dd.ANNOT != "INDEL"], dd[dd.ANNOT == "INDEL"].drop_duplicates(
        ["CHROM", "START", "END", "REF", "ALT", "GENE"])])
    dd = dd[~dd.ANNOT.isin(["Noncoding", "Synonymous", "Essential_Splice"]) & dd.GENE.isin(panel)]
    cj = len(dd) / 1234
    gm = storage.Store(p("pretrained"), "r").read_table("genic_model")
    gm = gm.set_index(gm.GENE).loc[panel]
    a_g, t_g = oracle.normal_params_to_gamma(gm.MU.values, gm.SIGMA.values)
    obs = oracle.gene_observed_counts(cds).reindex(panel).fillna(0)
    pis = {"SYN": gm.P_SILENT.values, "MIS": gm.P_MIS.values, "TRUNC": gm.P_NONS.values + gm.P_SPLICE.values}
    ks = {"SYN": obs.OBS_SYN.values.astype(float), "MIS": obs.OBS_MIS.values.astype(float),
          "TRUNC": (obs.OBS_NONS + obs.OBS_SPL).values.astype(float)}
    for c in ("SYN", "MIS", "TRUNC"):
        e_, p_ = oracle.burden_test(ks[c], a_g, t_g * cj, pis[c])
        assert np.array_equal(res["OBS_" + c].values, ks[c]), c
        np.testing.assert_allclose(res["EXP_" + c].values, e_, rtol=1e-9)
        assert_pvals_close(res["PVAL_%s_BURDEN" % c].values, p_)


def test_tiled_model_cli(workdir, oracle):
    """DigPreprocess.py preprocess_tiled + DigPretrain.py tiledModel (tiled_nonc_model, genic_driver_tools.py:599-690):
    every tile takes the parameters of the ONE window containing its start.  The flow project uses 1 kb windows, and
    the reference's lookup is hard-wired to a 10 kb grid (:636) -> KeyError there; a 10 kb project checks the numbers."""
    from digdriver_b200 import storage
    from digdriver_b200.sequence_model import genic_driver_tools as gd
    d = workdir["dir"]
    p = lambda x: str(d / x)
    if not storage.Store(p("pretrained"), "r").has("sequence_model_192") or not os.path.isdir(p("eltdata")):
        test_cli_flow_matches_oracle(workdir, oracle)
    rng = np.random.default_rng(31)
    tiles = []
    for c, L in ((1, 60_000), (2, 45_500)):
        for s in range(2000, L - 4000, 1700):
            tiles.append((c, s, s + int(rng.integers(50, 900)), "t%d_%d" % (c, s), 0, "+"))
    pd.DataFrame(tiles).to_csv(p("tiles.bed"), sep="\t", header=False, index=False)
    _cli("DigPreprocess", "preprocess_tiled %s %s %s 1000 TILES" % (p("tiles.bed"), p("eltdata"), p("genome.fa")))
    with pytest.raises(KeyError):
        _cli("DigPretrain", "tiledModel %s %s TILES" % (p("pretrained"), p("eltdata")))
    # ---- a 10 kb region model over the same genome: numbers against the oracle's element transfer
    W10 = 10_000
    wins = np.array([(c, i, i + W10) for c, L in ((1, 60_000), (2, 45_500)) for i in range(0, L - W10, W10)])
    rp = pd.DataFrame({"CHROM": wins[:, 0], "START": wins[:, 1], "END": wins[:, 2], "Y_TRUE": rng.poisson(20, len(wins)),
                       "Y_PRED": rng.gamma(2.0, 10.0, len(wins)), "STD": rng.uniform(0.5, 5.0, len(wins)),
                       "FLAG": rng.random(len(wins)) < 0.3}, index=["chr%d:%d-%d" % tuple(r) for r in wins])
    off2 = 60_032
    seq = np.full(off2 + 45_500, ord("N"), dtype=np.uint8)
    seq[:60_000] = workdir["seqs"]["chr1"]
    seq[off2:] = workdir["seqs"]["chr2"]
    off, ln = np.array([0, off2]), np.array([60_000, 45_500])
    wc, _ = oracle.count_regions(seq, off, ln, wins[:, 0] - 1, wins[:, 1], wins[:, 2], 1, 1)
    d_pr = np.random.default_rng(1).lognormal(np.log(1e-6), 1.0, 192)
    L_table = storage.Store(p("eltdata"), "r").read_table("TILES/L_counts")
    have = {(int(c), int(s)) for c, s, _ in wins}
    keep = [t for t in L_table.index
            if (int(t.split(":")[0][3:]), int(t.split(":")[1].split("-")[0]) // W10 * W10) in have]
    assert len(keep) > 40
    got = gd.tiled_model_arrays(keep, L_table, gd.RegionModel(rp), wc.astype(np.int32), d_pr)
    chrom = np.array([int(t.split(":")[0][3:]) for t in keep])
    start = np.array([int(t.split(":")[1].split("-")[0]) for t in keep])
    win_index = {(int(c), int(s)): i for i, (c, s, e) in enumerate(wins)}
    want = oracle.element_transfer(chrom, np.ones(len(keep), dtype=np.int8), np.arange(len(keep) + 1), start, start + 1,
                                   L_table.loc[keep].values, W10, win_index, wc, rp.Y_PRED.values, rp.STD.values,
                                   rp.Y_TRUE.values.astype(float), rp.FLAG.values, d_pr)
    assert list(got.ELT) == ["region_%d_%s" % (c, t.split(":")[1].replace("-", "_")) for c, t in zip(chrom, keep)]
    for col in ("MU", "SIGMA", "P_SUM"):
        np.testing.assert_allclose(got[col].values, want[col], rtol=1e-9, err_msg=col)
    assert np.array_equal(got.R_SIZE.values, want["R_SIZE"]) and np.array_equal(got.FLAG.values, want["FLAG"])
    assert np.array_equal(got.ELT_SIZE.values, (L_table.loc[keep].values.sum(axis=1) / 3).astype(np.int64))
    # the window really is the one containing the tile's start
    r = wc[[win_index[(c, s // W10 * W10)] for c, s in zip(chrom, start)]].sum(axis=1)
    assert np.array_equal(got.R_SIZE.values, r)


def _ref_region_counts_192(win_counts, overlaps, win_index, minus):
    """The reference's own expressions for region_counts (sequence_tools.py:625-634), restated: np.repeat of every
    overlapped window's 64 counts summed, re-ordered through the reverse-complement substitution names for '-'."""
    from digdriver_b200.sequence_model import sequence_tools as st
    subst_idx = st.mk_trans_idx(1, 1)
    revc = {s: st.reverse_complement(s.split('>')[0]) + '>' + st.reverse_complement(s.split('>')[-1]) for s in subst_idx}
    rc = np.array([np.repeat(win_counts[win_index[(int(c), int(s))]], 3) for c, s, e in overlaps]).sum(axis=0)
    if minus:
        rc = np.array([r[1] for r in sorted(enumerate(rc), key=lambda k: revc[subst_idx[k[0]]])])
    return rc


def test_persisted_intermediates_and_chunk_workers(workdir, oracle, tmp_path):
    """preprocess_nonc / preprocess_sites persist L_counts, region_counts and overlaps per element; nonc_model,
    genic_model and tiled_nonc_model (the reference's pool workers) read them back for a chunk of names."""
    from digdriver_b200 import storage
    from digdriver_b200.data_tools import mutation_tools
    from digdriver_b200.sequence_model import genic_driver_tools as gd, sequence_tools as st
    d = workdir["dir"]
    p = lambda x: str(d / x)
    if not storage.Store(p("pretrained"), "r").has("K1"):
        test_cli_flow_matches_oracle(workdir, oracle)
    pre = storage.Store(p("pretrained"), "r")
    em = pre.read_table("K1")
    genes, wins = workdir["genes"], workdir["wins"]
    # ---- preprocess_nonc with the reference's signature, into the same element data store
    L_contexts = st.precount_region_contexts_parallel(p("elements.bed"), p("genome.fa"), 1, W)
    st.preprocess_nonc(p("elements.bed"), p("eltdata"), p("pretrained"), L_contexts, "K2", W)
    data = storage.Store(p("eltdata"), "r")
    names, L, R, overlaps = data.read_element_groups("window_%d/K2" % W)
    assert names == [g[0] for g in genes]
    win_counts = data.read_array("window_%d/full_window_si_values" % W)
    win_index = {(int(c), int(s)): i for i, (c, s, e) in enumerate(data.read_array("window_%d/full_window_si_index" % W))}
    for i, g in enumerate(genes):
        ov = gd.get_ideal_overlaps(g[1], np.vstack((g[3], g[4])), W)
        assert overlaps[i] == [(int(c), int(s), int(e)) for c, s, e in ov]
        assert np.array_equal(R[i], _ref_region_counts_192(win_counts, ov, win_index, g[2] == "-")), g[0]
        want_L = sum(L_contexts.loc["chr%d:%d-%d" % (g[1], s, e)].values for s, e in zip(g[3], g[4]))
        assert np.array_equal(L[i], want_L)
    # ---- nonc_model on a chunk (reversed order) == the rows of the all-in-one element model, bit for bit
    chunk = [g[0] for g in genes][::-1][:17]
    got = gd.nonc_model(chunk, p("pretrained"), p("eltdata"), "K2", False)
    want = em.set_index("ELT").loc[chunk]
    assert list(got.ELT) == chunk and list(got.columns) == list(em.columns)
    for col in got.columns[1:]:
        assert np.array_equal(got[col].values, want[col].values, equal_nan=True), col
    with pytest.raises(KeyError):
        gd.nonc_model(["no_such_element"], p("pretrained"), p("eltdata"), "K2", False)
    # ---- indels_direct: a second region-parameter table for the indel columns
    rp2 = workdir["rp"].copy()
    rp2["Y_PRED"] = rp2.Y_PRED * 0.25
    rp2["Y_TRUE"] = rp2.Y_TRUE + 3
    pre_dir = tmp_path / "pre2"
    st2 = storage.Store(str(pre_dir), "w")
    st2.write_table("region_params", workdir["rp"])
    st2.write_table("region_params_indels", rp2)
    st2.write_table("sequence_model_192", pre.read_table("sequence_model_192"))
    gi = gd.nonc_model(chunk, str(pre_dir), p("eltdata"), "K2", True)
    np.testing.assert_allclose(gi.MU_INDEL.values, 0.25 * gi.MU.values, rtol=1e-12)
    assert np.array_equal(gi.R_INDEL.values, gi.R_OBS.values + 3 * np.array([len(overlaps[names.index(n)]) for n in chunk]))
    assert np.array_equal(gi.SIGMA_INDEL.values, gi.SIGMA.values)
    # the all-in-one element model honours the flag too (same rows as the chunk worker), and without it copies MU
    full = gd.nonc_model_parallel(str(pre_dir), p("eltdata"), "K1", indels_direct=True).set_index("ELT").loc[chunk]
    for col in ("MU", "SIGMA", "R_OBS", "MU_INDEL", "SIGMA_INDEL", "R_INDEL", "P_SUM"):
        assert np.array_equal(full[col].values, gi[col].values), col
    plain = gd.nonc_model_parallel(str(pre_dir), p("eltdata"), "K1", indels_direct=False)
    assert np.array_equal(plain.MU_INDEL.values, plain.MU.values) and np.array_equal(plain.R_INDEL.values, plain.R_OBS.values)
    # ---- preprocess_sites + nonc_model == sites_model_arrays
    ann = pd.read_table(p("annot.tsv"), header=None)
    ann = ann[ann[7] != "INDEL"].iloc[:600].copy()
    rng = np.random.default_rng(8)
    ann[5] = ["SET%02d" % s for s in rng.integers(0, 23, len(ann))]
    set_chrom = {s: int(c) for s, c in zip(ann[5], ann[0])}                        # one chromosome / strand per set
    ann = ann[[set_chrom[s] == int(c) for s, c in zip(ann[5], ann[0])]]
    ann[10] = ["-" if int(s[3:]) % 3 == 0 else "+" for s in ann[5]]
    ann.to_csv(p("sites.tsv"), sep="\t", header=False, index=False)
    st.preprocess_sites(p("sites.tsv"), p("eltdata"), p("pretrained"), "S1", W)
    set_names = sorted(set(ann[5]))
    got_s = gd.nonc_model(set_names, p("pretrained"), p("eltdata"), "S1", False)
    rm = gd.RegionModel(pre.read_table("region_params"))
    d_pr = st.d_pr_from_model192(pre.read_table("sequence_model_192"))
    key = {tuple(r): i for i, r in enumerate(map(tuple, data.read_array("window_%d/full_window_si_index" % W)))}
    rows = np.array([key[(int(c), int(s), int(e))] for c, s, e in zip(rm.df.CHROM, rm.df.START, rm.df.END)])
    want_s = gd.sites_model_arrays(mutation_tools.read_mutation_file(p("sites.tsv")), rm,
                                   win_counts[rows].astype(np.int32), d_pr)
    assert list(got_s.ELT) == list(want_s.ELT)
    for col in got_s.columns[1:]:
        assert np.array_equal(got_s[col].values, want_s[col].values, equal_nan=True), col
    # ---- genic_model on a chunk == rows of genic_model_parallel
    gm = pre.read_table("genic_model")
    gchunk = [g[0] for g in genes][5:25:3]
    gg = gd.genic_model(gchunk, p("pretrained"), p("genic"), "window_10kb/counts", False, f_fasta=p("genome.fa"))
    wantg = gm.set_index("GENE").loc[gchunk]
    assert list(gg.GENE) == gchunk
    for col in ("MU", "SIGMA", "R_OBS", "R_SIZE", "GENE_LENGTH", "P_MIS", "P_NONS", "P_SILENT", "P_SPLICE", "P_INDEL"):
        assert np.array_equal(gg[col].values, wantg[col].values), col
    # ---- nonc_model_region on bed12 rows == the element model (root-level window keys, as the reference reads them)
    rdir = tmp_path / "regiondata"
    rs = storage.Store(str(rdir), "w")
    rs.write_array("full_window_si_index", data.read_array("window_%d/full_window_si_index" % W))
    rs.write_array("full_window_si_values", win_counts)
    rs.write_table("Lkey", L_contexts)
    res, df_L, df_pi = gd.nonc_model_region(mutation_tools.bed12_boundaries(p("elements.bed")), p("pretrained"), str(rdir),
                                            "Lkey", return_intermediates=True)
    for col in ("R_OBS", "MU", "SIGMA", "P_SUM"):
        assert np.array_equal(res[col].values, em[col].values), col
    assert np.array_equal(df_L.values, L)
    np.testing.assert_allclose((df_pi.values * df_L.values).sum(axis=1), em.P_SUM.values, rtol=1e-12)
    par = gd.nonc_model_region_parallel(p("elements.bed"), p("pretrained"), str(rdir), "Lkey", 2)
    assert np.array_equal(par.P_SUM.values, em.P_SUM.values)


def test_si_by_regions_and_fetch_sequence(workdir, oracle):
    from digdriver_b200.genome import Genome
    from digdriver_b200.sequence_model import sequence_tools as st
    d = workdir["dir"]
    fa = str(d / "genome.fa")
    trans_idx = st.mk_trans_idx(1, 1)
    regions = ["chr1:0-1000", "chr1:3000-4000", "chr2:44000-45500"]
    off2 = 60_032
    seq = np.full(off2 + 45_500, ord("N"), dtype=np.uint8)
    seq[:60_000] = workdir["seqs"]["chr1"]
    seq[off2:] = workdir["seqs"]["chr2"]
    off, ln = np.array([0, off2]), np.array([60_000, 45_500])
    names = oracle.context_names(1, 1)
    for strand in (1, -1, '-'):
        got = st.si_by_regions(fa, trans_idx, regions, strand=strand)
        sc = np.full(3, -1 if strand in (-1, '-') else 1, dtype=np.int8)
        c64, _ = oracle.count_regions(seq, off, ln, np.array([0, 0, 1]), np.array([0, 3000, 44000]),
                                      np.array([1000, 4000, 45500]), 1, 1, strand=sc)
        tot = c64.sum(axis=0)
        assert list(got.index) == trans_idx
        assert np.array_equal(got[0].values, [tot[names.index(t.split('>')[0])] for t in trans_idx])
    assert st.si_by_regions(fa, trans_idx, []).values.sum() == 0
    g = Genome.from_fasta(fa)
    s, a, b = st.fetch_sequence(g, "chr1", 0, 10, n_up=2, n_down=2)
    assert (a, b) == (0, 12) and s == g.fetch("chr1", 0, 12).upper()
    s, a, b = st.fetch_sequence(fa, "chr2", 100, 110, n_up=1, n_down=1)
    assert (a, b) == (99, 111) and s == g.fetch("chr2", 99, 111).upper() and len(s) == 12
    with pytest.raises(ValueError):
        st.fetch_sequence(g, "chr1", 1, 10, n_up=2, n_down=2)


def test_count_mutations_and_genic_precount_cli(workdir, oracle, tmp_path, monkeypatch):
    """DigPretrain.py countMutations (cohort-size attributes) and DigPreprocess.py preprocess_genic_model
    (si_count_parallel: 192-substitution counts of the windows each gene overlaps)."""
    from digdriver_b200 import storage
    from digdriver_b200.sequence_model import genic_driver_tools as gd, sequence_tools as st
    d = workdir["dir"]
    p = lambda x: str(d / x)
    if not os.path.exists(p("annot.tsv")):
        test_cli_flow_matches_oracle(workdir, oracle)
    genes = workdir["genes"]
    (tmp_path / "genes_MSK_230.txt").write_text("\n".join(g[0] for g in genes[:12]) + "\n")
    monkeypatch.setenv("DIG_DATA_DIR", str(tmp_path))
    _cli("DigPretrain", "countMutations --outputFile %s --mutation-file %s" % (p("pretrained"), p("annot.tsv")))
    attrs = storage.Store(p("pretrained"), "r").get_attrs()
    ann = pd.read_table(p("annot.tsv"), header=None).drop_duplicates([0, 1, 2, 3, 4, 5])
    snv, ind = ann[ann[7] != "INDEL"], ann[ann[7] == "INDEL"].drop_duplicates([0, 1, 2, 3, 4, 6])
    dd = pd.concat([snv, ind])
    cds = dd[dd[7] != "Noncoding"]
    rp = workdir["rp"]
    assert attrs["N_SAMPLES"] == dd[5].nunique() and attrs["N_MUT_CDS"] == len(cds) == attrs["N_MUT_SAMPLE_CDS"]
    assert attrs["N_MUT_TOTAL"] == rp.Y_TRUE.sum() and attrs["N_MUT_TRAIN"] == rp.Y_TRUE[~rp.FLAG].sum()
    sel = cds[cds[6].isin([g[0] for g in genes[:12]]) & ~cds[7].isin(["Synonymous", "Essential_Splice", "Noncoding"])]
    assert attrs["N_MUT_MSK_230"] == len(sel) > 0
    assert attrs["N_MUT_SAMPLE_MSK_230"] == len(sel.drop_duplicates([5, 6])) and attrs["N_SAMPLE_MSK_230"] == sel[5].nunique()
    # ---- preprocess_genic_model
    _cli("DigPreprocess", "preprocess_genic_model %s %s %s --out-key cds/w --window %d" % (p("genic"), p("genome.fa"), p("si"), W))
    si = storage.Store(p("si"), "r").read_table("cds/w")
    assert list(si.index) == [g[0] for g in genes] and list(si.columns) == st.mk_trans_idx(1, 1)
    cs = storage.Store(p("counts"), "r")
    wc = cs.read_table("all_window_genome_counts")
    for g in genes[:8]:
        ov = gd.get_ideal_overlaps(str(g[1]), np.vstack((g[3], g[4] - 1)), W)
        tot = wc.loc[[gd.trip_to_str(r) for r in ov]].values.sum(axis=0)
        if g[2] == "-":
            tot = tot[[list(wc.columns).index(st.reverse_complement(c)) for c in wc.columns]]
        assert np.array_equal(si.loc[g[0]].values, np.repeat(tot, 3)), g[0]
