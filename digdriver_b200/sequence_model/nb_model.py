"""Drop-in for the p-value arithmetic of DIGDriver/sequence_model/nb_model.py.  The SciPy calls
(scipy.special.betainc, scipy.stats.nbinom.pmf) are replaced by the FP64 kernel K7."""
import numpy as np
import torch

from .. import kernels


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


def normal_params_to_gamma(mu, sigma):
    """Reference :237-241."""
    alpha = mu ** 2 / sigma ** 2
    theta = sigma ** 2 / mu
    return alpha, theta


def _like(ref, values):
    """Return ``values`` as the same kind of object the reference would (Series in -> Series out)."""
    try:
        import pandas as pd
        if isinstance(ref, pd.Series):
            return pd.Series(values, index=ref.index)
    except Exception:
        pass
    if np.ndim(ref) == 0:
        return float(values[0])
    return values


def nb_pvalue_greater_midp(k, alpha, p):
    """UPPER TAIL p-value of a negative binomial with a mid-p correction (reference :271-278):
    0.5 * nbinom.pmf(k, alpha, p) + betainc(k + 1, alpha, 1 - p), evaluated by the GPU kernel."""
    kb, ab, pb = np.broadcast_arrays(np.asarray(k, dtype=np.float64), np.asarray(alpha, dtype=np.float64),
                                     np.asarray(p, dtype=np.float64))
    shape = kb.shape
    out = kernels.nb_pvalue_greater_midp(np.ascontiguousarray(kb).reshape(-1), np.ascontiguousarray(ab).reshape(-1),
                                         np.ascontiguousarray(pb).reshape(-1), _dev()).cpu().numpy()
    for ref in (k, alpha, p):
        if hasattr(ref, "index"):
            return _like(ref, out)
    return out.reshape(shape) if shape else float(out[0])


def get_q_vals(pvals_lst):
    """Benjamini-Hochberg FDR (statsmodels fdrcorrection, method='indep') of reference :340-342."""
    p = np.asarray(pvals_lst, dtype=np.float64)
    n = len(p)
    order = np.argsort(p)
    ranked = p[order] * n / np.arange(1, n + 1)
    q = np.minimum.accumulate(ranked[::-1])[::-1]
    q = np.minimum(q, 1.0)
    out = np.empty(n)
    out[order] = q
    return out


def nb_pvalue_exact(k, alpha, p, mu=None):
    """UPPER or LOWER TAIL p-value of a negative binomial, chosen by whether k is below the expectation
    (reference :298-314), evaluated by the GPU kernel K8.  ``mu`` (the reference's optional override of the
    expectation alpha (1-p)/p) is accepted for signature parity; a truthy value other than that expectation is not
    supported."""
    if mu:
        raise NotImplementedError("nb_pvalue_exact: only the default expectation alpha*(1-p)/p is supported")
    kb, ab, pb = np.broadcast_arrays(np.asarray(k, dtype=np.float64), np.asarray(alpha, dtype=np.float64),
                                     np.asarray(p, dtype=np.float64))
    shape = kb.shape
    out = kernels.nb_pvalue_exact(np.ascontiguousarray(kb).reshape(-1), np.ascontiguousarray(ab).reshape(-1),
                                  np.ascontiguousarray(pb).reshape(-1), _dev()).cpu().numpy()
    return out.reshape(shape) if shape else float(out[0])


def _mutation_starts(tabix):
    """The reference reads mutations through pysam.TabixFile (absent here): accept a DataFrame with CHROM / START
    columns, a (chrom, start) array pair, or a path to a mutation file readable by data_tools.read_mutation_file."""
    import pandas as pd
    if isinstance(tabix, (str, bytes)):
        from ..data_tools import mutation_tools
        tabix = mutation_tools.read_mutation_file(tabix, drop_sex=False)
    if isinstance(tabix, pd.DataFrame):
        return tabix.CHROM.astype(str).str.replace("^chr", "", regex=True).values, tabix.START.values.astype(np.int64)
    chrom, start = tabix
    return np.asarray(chrom).astype(str), np.asarray(start, dtype=np.int64)


def nb_model(d_pr, idx, mu_lst, sigma_lst, f_tabix, f_fasta, n_up=2, n_down=2, binsize=50, collapse=False):
    """Per-position (binsize=1) or per-bin NB hotspot test over the regions ``idx`` = rows of (CHROM, START, END)
    (reference :188-235): one K8 launch for all regions instead of a Python loop per base.  Returns the reference's
    DataFrame (CHROM, POS, OBS, EXP, PVAL, Pi, MU, SIGMA, REGION)."""
    import pandas as pd
    from . import sequence_tools
    g = sequence_tools.get_device_genome(f_fasta)
    idx = np.asarray(idx)
    chroms = [str(c) for c in idx[:, 0]]
    starts, ends = idx[:, 1].astype(np.int64), idx[:, 2].astype(np.int64)
    if np.any((starts > 0) & (starts < n_up)):
        raise ValueError("start out of range")
    cidx = g.chrom_indices(["chr{}".format(c) for c in chroms], prefix="")
    m_chrom, m_start = _mutation_starts(f_tabix)
    name_to_idx = {str(n).replace("chr", "", 1): i for i, n in enumerate(g.names)}
    keep = np.array([c in name_to_idx for c in m_chrom], dtype=bool)
    m_cidx = np.array([name_to_idx[c] for c in m_chrom[keep]], dtype=np.int32)
    mu = np.asarray(mu_lst, dtype=np.float64)
    sigma = np.asarray(sigma_lst, dtype=np.float64)
    out = kernels.position_test(g, cidx, starts, ends, mu, sigma,
                                sequence_tools._s_prob_table(d_pr, n_up, n_down, collapse), m_cidx, m_start[keep],
                                n_up=n_up, n_down=n_down, binsize=binsize)
    nb = np.diff(out["bin_ptr"])
    rep = lambda v: np.repeat(np.asarray(v), nb)
    df = pd.DataFrame({"CHROM": rep([int(c) if c.isdigit() else c for c in chroms]),
                       "POS": out["pos"].cpu().numpy(), "OBS": out["obs"].cpu().numpy().astype(np.float64),
                       "EXP": out["exp"].cpu().numpy(), "PVAL": out["pval"].cpu().numpy(),
                       "Pi": out["pt"].cpu().numpy(), "MU": rep(mu), "SIGMA": rep(sigma)})
    df["REGION"] = rep(["{}:{}-{}".format(c, s, e) for c, s, e in zip(chroms, starts, ends)])
    return df


def apply_nb_to_region(CHROM, START, END, mu, sigma, S_probs, tabix, fasta, n_up=2, n_down=2, binsize=1,
                       collapse=False):
    """One region of nb_model (reference :126-186): returns (pvals, poss, obss, exps, pt_lst)."""
    df = nb_model(S_probs, np.array([[CHROM, START, END]], dtype=object), [mu], [sigma], tabix, fasta, n_up=n_up,
                  n_down=n_down, binsize=binsize, collapse=collapse)
    return (df.PVAL.values, df.POS.values, df.OBS.values.astype(np.int64), df.EXP.values, list(df.Pi.values))
