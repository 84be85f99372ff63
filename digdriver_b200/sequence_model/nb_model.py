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
