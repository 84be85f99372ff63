#!/usr/bin/env python
"""Golden vectors for the secondary gene tests from the UNMODIFIED reference functions
transfer_tools.gene_expected_muts_dnds / gene_pvalue_burden_dnds / gene_pvalue_sel_nb / selection_coefficient.

    python tests/golden/make_golden_secondary.py        (build container only)
"""
import os
import sys
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402


def main():
    tt = rh.load_reference().transfer_tools
    rng = np.random.default_rng(777)
    n = 600
    mu = rng.gamma(2.0, 4.0, n) + 0.01
    sigma = mu * rng.uniform(0.05, 1.6, n)                      # sigma > mu gives ALPHA < 1 (the _mle_t branch)
    alpha, theta = mu ** 2 / sigma ** 2, sigma ** 2 / mu
    theta = theta * rng.uniform(0.2, 3.0)                        # a scale factor
    pi = rng.dirichlet([2.3, 6.8, 0.4, 0.5], n) * rng.uniform(0.001, 0.05, n)[:, None]
    pi[::37, 2] = 0.0                                            # a class with zero target size
    lam = (alpha * theta)[:, None] * pi
    obs = rng.poisson(lam * rng.choice([0.3, 1.0, 1.0, 4.0, 25.0], (n, 4)))
    obs[::11] = 0
    df = pd.DataFrame({"ALPHA": alpha, "THETA": theta, "Pi_SYN": pi[:, 0], "Pi_MIS": pi[:, 1], "Pi_NONS": pi[:, 2],
                       "Pi_SPL": pi[:, 3], "OBS_SYN": obs[:, 0], "OBS_MIS": obs[:, 1], "OBS_NONS": obs[:, 2],
                       "OBS_SPL": obs[:, 3]}, index=["G%d" % i for i in range(n)])
    df["Pi_TRUNC"] = df.Pi_NONS + df.Pi_SPL
    df["Pi_NONSYN"] = df.Pi_MIS + df.Pi_TRUNC
    df["OBS_TRUNC"] = df.OBS_NONS + df.OBS_SPL
    df["OBS_NONSYN"] = df.OBS_MIS + df.OBS_TRUNC
    inputs = df.copy()
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        df = tt.gene_expected_muts_dnds(df)
        df = tt.gene_pvalue_burden_dnds(df)
        df = tt.gene_pvalue_sel_nb(df)
        for c in ("SYN", "MIS", "TRUNC"):
            tt.selection_coefficient(df, c, pvalue=True)
    out = {"in_" + c: inputs[c].values.astype(np.float64) for c in inputs.columns}
    for c in df.columns:
        if c not in inputs.columns:
            out["out_" + c] = np.asarray(df[c].values, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "secondary.npz"), **out)
    print("wrote secondary.npz:", [c for c in df.columns if c not in inputs.columns])


if __name__ == "__main__":
    main()
