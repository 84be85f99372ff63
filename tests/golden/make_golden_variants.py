#!/usr/bin/env python
"""Golden vectors for the remaining p-value conventions of nb_model.py and the remaining selection / indel tests of
transfer_tools.py, from the UNMODIFIED reference functions (scalar / row loops, exactly as the reference calls them):

    nb_model.nb_pvalue_greater, nb_pvalue_greater_midp_DEPRECATED, nb_pvalue_less_midp, nb_pvalue_midp (with and
    without mu), nb_pvalue_exact (with mu); nb_pvalue_less returns None in the reference, so its golden is the
    expression the function evaluates (scipy.special.betainc(alpha, k+1, p), nb_model.py:283).
    transfer_tools._ll_nb / _ll_pois / _ll_gamma, _llr_test_nb, _llr_test_gamma_poiss, gene_pvalue_sel_gamma,
    gene_pvalue_indel_by_transfer (package data redirected to a small generated gene table), _mle_t, _mrfold_factor.

    python tests/golden/make_golden_variants.py        (build container only)
"""
import gzip
import os
import sys
import tempfile
import warnings

import numpy as np
import pandas as pd
import scipy.special

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402


def main():
    ref = rh.load_reference()
    nb, tt = ref.nb_model, ref.transfer_tools
    out = {}
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        # ---- p-value conventions on the grid of nbtest.npz (valid parameter domain) + k = 0 / p = 1 edges
        z = np.load(os.path.join(HERE, "nbtest.npz"))
        k, a, p = z["k"], z["alpha"], z["p"]
        ok = np.isfinite(k) & np.isfinite(a) & np.isfinite(p) & (k >= 0) & (a > 0) & (p > 0) & (p <= 1) & (k == np.floor(k))
        k, a, p = k[ok], a[ok], p[ok]
        rng = np.random.default_rng(4242)
        ke = np.concatenate([np.zeros(40), rng.integers(0, 50, 60).astype(float), np.array([0., 1., 7.])])
        ae = np.concatenate([rng.gamma(2.0, 3.0, 100) + 1e-3, np.array([2.5, 2.5, 2.5])])
        pe = np.concatenate([rng.uniform(1e-4, 1.0, 100), np.array([1.0, 1.0, 1.0])])
        k, a, p = np.concatenate([k, ke]), np.concatenate([a, ae]), np.concatenate([p, pe])
        mean = a * (1 - p) / p
        mu = mean * rng.choice([0.25, 0.9, 1.0, 1.1, 4.0], len(k))
        mu[::9] = 0.0                                                   # falsy -> the default expectation
        out.update(v_k=k, v_alpha=a, v_p=p, v_mu=mu)
        f = lambda fn, *cols: np.array([fn(*[float(c[i]) for c in cols]) for i in range(len(k))], dtype=np.float64)  # noqa: E731
        out["v_greater"] = f(nb.nb_pvalue_greater, k, a, p)
        out["v_greater_midp_deprecated"] = f(nb.nb_pvalue_greater_midp_DEPRECATED, k, a, p)
        out["v_less"] = f(lambda kk, aa, pp: scipy.special.betainc(aa, kk + 1, pp), k, a, p)
        out["v_less_midp"] = f(nb.nb_pvalue_less_midp, k, a, p)
        out["v_midp"] = f(nb.nb_pvalue_midp, k, a, p)
        out["v_midp_mu"] = f(lambda kk, aa, pp, mm: nb.nb_pvalue_midp(kk, aa, pp, mu=mm), k, a, p, mu)
        out["v_exact_mu"] = f(lambda kk, aa, pp, mm: nb.nb_pvalue_exact(kk, aa, pp, mu=mm), k, a, p, mu)

        # ---- log-likelihood terms and row-level LLR tests on the rows of secondary.npz
        s = np.load(os.path.join(HERE, "secondary.npz"))
        df = pd.DataFrame({c[3:]: s[c] for c in s.files if c.startswith("in_")})
        for c in ("T_SYN", "MRFOLD", "EXP_SYN"):
            df[c] = s["out_" + c]
        out["ll_nb"] = np.asarray(tt._ll_nb(df.OBS_MIS.values, df.ALPHA.values, (df.THETA * df.Pi_MIS).values))
        out["ll_pois"] = np.asarray(tt._ll_pois(df.OBS_MIS.values, (df.ALPHA * df.THETA * df.Pi_MIS).values))
        out["ll_pois_self"] = np.asarray(tt._ll_pois(df.OBS_NONS.values, df.OBS_NONS.values))
        out["ll_gamma"] = np.asarray(tt._ll_gamma(df.T_SYN.values, df.ALPHA.values, (df.THETA * df.Pi_SYN * df.MRFOLD).values))
        out["llr_nb_rows"] = np.array([tt._llr_test_nb(row) for _, row in df.iterrows()], dtype=np.float64)
        out["llr_pg_rows"] = np.array([tt._llr_test_gamma_poiss(row) for _, row in df.iterrows()], dtype=np.float64)
        dfg = tt.gene_pvalue_sel_gamma(df.copy())
        for c in ("SYN", "MIS", "NONS", "NONSYN"):
            out["PVAL_%s_SEL_PG" % c] = dfg["PVAL_%s_SEL_PG" % c].values.astype(np.float64)
        out["mle_t"] = np.array([tt._mle_t(r.OBS_SYN, 1, r.ALPHA, r.THETA * r.Pi_SYN) for _, r in df.iterrows()])
        out["mrfold"] = np.array([tt._mrfold_factor(r.T_SYN, r.EXP_SYN) for _, r in df.iterrows()])

        # ---- gene_pvalue_indel_by_transfer with the package data redirected to a generated gene table
        n = 300
        genes = ["GENE%03d" % i for i in range(n)]
        tmp = tempfile.mkdtemp()
        os.makedirs(os.path.join(tmp, "data"))
        rows, lengths = [], np.zeros(n, dtype=np.int64)
        for i, g in enumerate(genes[:-5]):                              # the last five genes have no CDS row (NaN)
            for b in range(int(rng.integers(1, 6))):
                st = int(rng.integers(1000, 10 ** 6))
                ln = int(rng.integers(50, 900))
                rows.append((str(1 + i % 22), st, st + ln, g))
                lengths[i] += ln
        with gzip.open(os.path.join(tmp, "data", "dndscv_gene_cds.bed.gz"), "wt") as fh:
            for r in rows:
                fh.write("\t".join(map(str, r)) + "\n")
        cgc = genes[3:40:4]
        with open(os.path.join(tmp, "data", "genes_CGC_ALL.txt"), "w") as fh:
            fh.write("\n".join(cgc) + "\n")
        pr = sys.modules["pkg_resources"]
        pr.resource_filename = lambda pkg, rel: os.path.join(tmp, rel)
        pr.resource_stream = lambda pkg, rel: open(os.path.join(tmp, rel), "rb")
        mu_g = rng.gamma(2.0, 4.0, n) + 0.01
        sg = mu_g * rng.uniform(0.05, 1.2, n)
        dfi = pd.DataFrame({"ALPHA": mu_g ** 2 / sg ** 2, "THETA": sg ** 2 / mu_g * 1.7,
                            "R_SIZE": rng.integers(20000, 200000, n).astype(np.int64),
                            "OBS_INDEL": rng.poisson(0.4, n).astype(np.float64)}, index=genes)
        dfo = tt.gene_pvalue_indel_by_transfer(dfi.copy())
        out.update(ind_genes=np.array(genes), ind_cds_chrom=np.array([r[0] for r in rows]),
                   ind_cds_start=np.array([r[1] for r in rows]), ind_cds_end=np.array([r[2] for r in rows]),
                   ind_cds_gene=np.array([r[3] for r in rows]), ind_cgc=np.array(cgc),
                   ind_ALPHA=dfi.ALPHA.values, ind_THETA=dfi.THETA.values, ind_R_SIZE=dfi.R_SIZE.values,
                   ind_OBS_INDEL=dfi.OBS_INDEL.values)
        for c in ("LENGTH", "Pi_INDEL", "THETA_INDEL", "EXP_INDEL", "PVAL_INDEL_BURDEN"):
            out["ind_out_" + c] = dfo[c].values.astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "variants.npz"), **out)
    print("wrote variants.npz:", {k2: v.shape for k2, v in out.items()})


if __name__ == "__main__":
    main()
