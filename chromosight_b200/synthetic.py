"""Seeded synthetic Hi-C contact maps.

The reference ships no generator; these maps stand in for the balanced
intra-chromosomal sub-matrices that ``ContactMap.create_mat``
(contacts_map.py:527-548) hands to the hot path: an upper band of Poisson
counts following a power-law distance decay, sparse at long range, with a few
percent of undetectable bins and planted loops.
"""
import numpy as np
import scipy.sparse as sp


def band_counts(n, n_diags, seed=0, missing_frac=0.02, loops_per_bin=0.01, max_dist=None):
    """Raw (not detrended) upper-band contact map.

    Parameters
    ----------
    n : int
        Number of bins of the chromosome.
    n_diags : int
        Diagonals 0..n_diags (inclusive) are populated.
    seed : int
        Seed of numpy.random.default_rng.
    missing_frac : float
        Fraction of undetectable bins; their rows and columns are zeroed.
    loops_per_bin : float
        Planted 3x3 blobs per bin.
    max_dist : int or None
        Loops are planted at distances U[12, max_dist - 10] (defaults to n_diags).

    Returns
    -------
    matrix : scipy.sparse.csr_matrix of float64, shape (n, n), upper triangle.
    detectable : numpy.ndarray of int, indices of detectable bins.
    """
    rng = np.random.default_rng(seed)
    W = n_diags + 1
    d = np.arange(W)
    lam = 200.0 / (1.0 + d) ** 0.8 + 2.0
    density = np.maximum(0.05, 1.0 / (1.0 + d / 50.0))
    band = rng.poisson(lam[None, :], size=(n, W)).astype(np.float64)
    band *= rng.random((n, W)) < density[None, :]
    if max_dist is None:
        max_dist = n_diags
    n_loops = int(n * loops_per_bin)
    if n_loops and max_dist > 24:
        lr = rng.integers(1, max(2, n - max_dist - 2), size=n_loops)
        ld = rng.integers(12, max_dist - 10, size=n_loops)
        for dr in (-1, 0, 1):
            for dc in (-1, 0, 1):
                rr = lr + dr
                dd = ld + dc - dr
                ok = (rr >= 0) & (rr < n) & (dd >= 0) & (dd < W)
                band[rr[ok], dd[ok]] = np.maximum(band[rr[ok], dd[ok]], 1.0) * 3.0 + 3.0
    missing = rng.random(n) < missing_frac
    rows = np.repeat(np.arange(n), W)
    cols = rows + np.tile(d, n)
    vals = band.ravel()
    keep = (cols < n) & (vals != 0)
    rows, cols, vals = rows[keep], cols[keep], vals[keep]
    keep = ~(missing[rows] | missing[cols])
    mat = sp.csr_matrix((vals[keep], (rows[keep], cols[keep])), shape=(n, n))
    return mat, np.flatnonzero(~missing)


def inter_counts(ms, ns, seed=0, density=1e-2, missing_frac=0.02):
    """Rectangular inter-chromosomal map: Gamma(2, 0.5) values at `density`."""
    rng = np.random.default_rng(seed)
    nnz = int(ms * ns * density)
    r = rng.integers(0, ms, size=nnz)
    c = rng.integers(0, ns, size=nnz)
    v = rng.gamma(2.0, 0.5, size=nnz)
    miss_r = rng.random(ms) < missing_frac
    miss_c = rng.random(ns) < missing_frac
    keep = ~(miss_r[r] | miss_c[c])
    mat = sp.coo_matrix((v[keep], (r[keep], c[keep])), shape=(ms, ns)).tocsr()
    mat.sum_duplicates()
    return mat, (np.flatnonzero(~miss_r), np.flatnonzero(~miss_c))


def n_windows(n, max_dist):
    """Pearson windows of an intra map scanned up to max_dist."""
    D = min(max_dist, n - 1)
    return (D + 1) * n - D * (D + 1) // 2
