"""Statistical helpers mirroring chromosight.utils.stats (stats.py:7-81)."""
import numpy as np
from scipy.special import erfc


def fdr_correction(pvals):
    """Benjamini-Hochberg q-values (stats.py:7-40); host side, the CLI applies it
    once on the global pattern table (cli:848)."""
    if pvals is None:
        return None
    p = np.asarray(pvals, dtype=float)
    n = len(p)
    order = np.argsort(p)[::-1]              # largest p first
    ranks = np.arange(n, 0, -1)              # rank of each sorted p-value
    q_sorted = np.minimum(1, np.minimum.accumulate(p[order] * n / ranks))
    q = np.empty(n)
    q[order] = q_sorted
    return q


def corr_to_pval(corr, n, rho0=0):
    """Two-sided log10 p-values of Pearson coefficients through Fisher's z
    (stats.py:43-81).  Host version for small arrays (final pattern tables);
    whole score maps get theirs on the GPU (csrc/scores.cu, log10_pval)."""
    corr = np.asarray(corr, dtype=float)
    if isinstance(n, np.ndarray):
        if n.shape != corr.shape:
            raise ValueError("corr and n must have identical shapes.")
    with np.errstate(all="ignore"):
        z = np.abs((np.arctanh(corr) - np.arctanh(rho0)) * np.sqrt(np.asarray(n, dtype=float) - 3))
        a = z * np.sqrt(0.5)
        p = np.where(a * a > 7.09782712893383996843e2, 0.0, erfc(a))
        p = np.where(np.isnan(z), np.nan, p)
        return np.log10(p)
