"""scipy.sparse port of chromosight's normalised cross-correlation -- the CPU
baseline arm of bench.py (`cpu_baseline.kind = "port"`, `--impl reference`).

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py); never imported by
the product package.  Parity: PINNED against the reference through
tests/golden/*.npz (tests/test_oracle_golden.py::test_sparse_port_*).

Unlike oracle/pearson_oracle.py (dense, per-window, an independent restatement
used as the checker), this module follows the reference's own evaluation
strategy so that timing it says something about the reference:

* every raw correlation is a sum of `km` products of a row-shifted slice of the
  sparse signal with a sparse Toeplitz matrix built from one kernel row
  (det:699-713), or two banded products for a constant kernel (det:682-692);
* the Pearson map is assembled from six such correlations (det:1002-1092) --
  mean of S, mean of S^2, S*K, and, on the mask, the count, K and K^2;
* only stored (non-zero) entries are ever touched.

The reference then patches the masked windows through ~9 sparse fancy-index
get/set passes; here the six maps are sampled once on the union of their
supports and combined with flat numpy arithmetic, which is a little faster than
the reference (measured in DESIGN.md) -- the baseline is not handicapped.
"""
import numpy as np
import scipy.sparse as sp

from .pearson_oracle import (DENOM_EPS, XCORR_THRESHOLD, corr_to_log10_pval,
                             truncate_kernel)


def _toeplitz_rows(values, n_out, n_in):
    """(n_out x n_in) sparse matrix T with T[r, r + j] = values[j]: one kernel
    row unrolled for a valid-mode 1-D correlation (det:704-713)."""
    k = len(values)
    return sp.diags(np.asarray(values, dtype=np.float64), np.arange(k), shape=(n_out, n_in),
                    format="csr")


def xcorr2_valid(signal, kernel):
    """Valid-mode cross-correlation of a CSR signal with a dense kernel as
    sparse products (det:627-713), not thresholded, shape (sm-km+1, sn-kn+1)."""
    signal = signal.tocsr()
    sm, sn = signal.shape
    km, kn = kernel.shape
    om, on = sm - km + 1, sn - kn + 1
    if np.allclose(kernel, kernel[0, 0], rtol=1e-8):
        # det:682-692: a constant kernel is two banded box sums
        left = _toeplitz_rows(np.full(km, kernel[0, 0]), om, sm)
        right = _toeplitz_rows(np.ones(kn), on, sn).T
        return ((left @ signal) @ right).tocsr()
    out = sp.csr_matrix((om, on), dtype=np.float64)
    if kn < km:
        # det:699-706: scan the short side
        for kj in range(kn):
            out = out + _toeplitz_rows(kernel[:, kj], om, sm) @ signal[:, kj:on + kj]
    else:
        for ki in range(km):
            out = out + signal[ki:om + ki, :] @ _toeplitz_rows(kernel[ki, :], on, sn).T
    return out.tocsr()


def xcorr2_sparse(signal, kernel, threshold=XCORR_THRESHOLD, tsvd=None):
    """det:595-624 for a sparse signal: thresholded correlation at the signal's
    shape (zero margins of half a kernel, det:720-722)."""
    kernel = np.asarray(kernel, dtype=np.float64)
    if tsvd is not None:
        kernel = truncate_kernel(kernel, tsvd)
    km, kn = kernel.shape
    out = xcorr2_valid(signal, kernel).tocoo()
    keep = np.abs(out.data) >= threshold                     # det:716
    kh, kw = (km - 1) // 2, (kn - 1) // 2
    return sp.csr_matrix((out.data[keep], (out.row[keep] + kh, out.col[keep] + kw)),
                         shape=signal.shape)


def _frame(signal, mk, nk):
    """det:979-985: zero margins of (mk-1, nk-1) around the signal."""
    coo = signal.tocoo()
    ms, ns = coo.shape
    return sp.csr_matrix((coo.data, (coo.row + mk - 1, coo.col + nk - 1)),
                         shape=(ms + 2 * (mk - 1), ns + 2 * (nk - 1)))


def _frame_mask(mask, kernel_shape, sym_upper, max_dist):
    """pre:404-498 on sparse matrices (same result as
    pearson_oracle.frame_missing_mask_dense)."""
    mk, nk = kernel_shape
    ms, ns = mask.shape
    coo = mask.tocoo()
    r, c = coo.row[coo.data != 0], coo.col[coo.data != 0]
    banded = sym_upper and max_dist is not None
    if banded:
        d = c - r
        ok = (d >= 0) & (d <= max_dist + max(nk, mk))       # pre:452-454
        r, c = r[ok], c[ok]
    H, W = ms + 2 * (mk - 1), ns + 2 * (nk - 1)
    rows, cols = [r + mk - 1], [c + nk - 1]

    def rect(y0, y1, x0, x1):
        if y1 > y0 and x1 > x0:
            yy, xx = np.meshgrid(np.arange(y0, y1), np.arange(x0, x1), indexing="ij")
            rows.append(yy.ravel())
            cols.append(xx.ravel())

    if banded:
        max_m, max_n = max_dist + mk, max_dist + nk
        rect(0, mk - 1, nk - 1, nk - 1 + min(max_n, ns))     # pre:461-463
        rect(max(0, H - (max_m + 1)), H, W - (nk - 1), W)    # pre:475
        rect(0, mk - 1, 0, nk - 1)                           # pre:477
    else:
        rect(0, mk - 1, 0, W)
        rect(H - (mk - 1), H, 0, W)
        rect(0, H, 0, nk - 1)
        rect(0, H, W - (nk - 1), W)
    if sym_upper:
        big_k = max(nk, mk)                                  # pre:483-497
        for off in range(1, big_k + 1):
            y = np.arange(off, min(H, W + off))
            rows.append(y)
            cols.append(y - off)
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    m = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(H, W))
    m.data[:] = 1.0
    return m


def _sample(mat, rows, cols):
    """Values of a CSR matrix at (rows, cols) -- csr_sample_values, as the
    reference's fancy indexing does (det:1037-1085)."""
    return np.asarray(mat[rows, cols]).ravel()


def normxcorr2_sparse(signal, kernel, max_dist=None, sym_upper=False, full=False,
                      missing_mask=None, missing_tol=0.75, tsvd=None, pval=False):
    """det:917-1131 on scipy.sparse matrices.  Returns (corr csr, log10 p csr or
    None) with the reference's sparsity (zeros eliminated)."""
    K = np.asarray(kernel, dtype=np.float64)
    mk, nk = K.shape
    N = mk * nk
    ones = np.ones((mk, nk))
    signal = signal.tocsr().astype(np.float64)
    if full:
        F = _frame(signal, mk, nk)
        M = None
        if missing_mask is not None:
            M = _frame_mask(missing_mask, K.shape, sym_upper, max_dist)
    else:
        F = signal
        M = None if missing_mask is None else missing_mask.tocsr().astype(np.float64)

    mean_s = xcorr2_sparse(F, ones / N)                       # det:1005 / 1050
    mean_s2 = xcorr2_sparse(F.power(2), ones / N)             # det:1007 / 1055
    mean_sk = xcorr2_sparse(F, K / N, tsvd=tsvd)              # det:1018 / 1082
    k_sum, k2_sum = K.sum(), (K ** 2).sum()
    k_mean, k2_mean = k_sum / N, k2_sum / N
    # union of the supports: pixels where any term can be non-zero
    support = (abs(mean_s) + abs(mean_s2) + abs(mean_sk)).tocoo()
    rows, cols = support.row, support.col
    a1, a2, a3 = _sample(mean_s, rows, cols), _sample(mean_s2, rows, cols), _sample(mean_sk, rows, cols)
    with np.errstate(all="ignore"):
        if M is None:
            denom = np.sqrt(a2 - a1 ** 2) * float(K.std())    # det:1007-1008
            num = a3 - a1 * float(K.mean())
            n_obs = np.full(len(rows), float(N))
        else:
            n_miss = _sample(xcorr2_sparse(M, ones), rows, cols)              # det:1030
            has = n_miss != 0
            n_pres = N - n_miss
            km_wm = (k_sum - _sample(xcorr2_sparse(M, K, tsvd=tsvd), rows, cols)) / n_pres        # det:1035
            k2m_wm = (k2_sum - _sample(xcorr2_sparse(M, K ** 2, tsvd=tsvd), rows, cols)) / n_pres  # det:1041
            scale = np.where(has, N / n_pres, 1.0)
            m_s, m_s2 = a1 * scale, a2 * scale
            var_k = np.where(has, k2m_wm - km_wm ** 2, k2_mean - k_mean ** 2)
            denom = np.sqrt((m_s2 - m_s ** 2) * var_k)        # det:1060-1066
            denom[has & (n_pres < int((1 - missing_tol) * N))] = 0.0          # det:1069-1072
            cov_miss = (a3 - a1 * km_wm) * scale
            if k_mean == 0:
                cov_miss = np.full(len(rows), np.nan)
            num = np.where(has, cov_miss, a3 - a1 * k_mean)   # det:1075-1085
            n_obs = np.where(has & (n_pres != 0), n_pres, float(N)) if full else np.full(len(rows), float(N))
        r = np.where(np.abs(denom) >= DENOM_EPS, num / denom, 0.0)            # det:1088-1092
    if sym_upper:
        r = np.where(cols >= rows, r, 0.0)                    # det:1098-1099
    r[~np.isfinite(r)] = 0.0
    r = np.clip(r, -1.0, 1.0)                                 # det:1105-1106
    keep = r != 0
    rows, cols, r, n_obs = rows[keep], cols[keep], r[keep], n_obs[keep]
    shape = F.shape
    corr = sp.csr_matrix((r, (rows, cols)), shape=shape)
    pvals = None
    if pval:
        with np.errstate(all="ignore"):
            pvals = sp.csr_matrix((corr_to_log10_pval(r, n_obs), (rows, cols)), shape=shape)
    if full:                                                  # det:1124-1129
        corr = corr[mk - 1:shape[0] - mk + 1, nk - 1:shape[1] - nk + 1]
        if pvals is not None:
            pvals = pvals[mk - 1:shape[0] - mk + 1, nk - 1:shape[1] - nk + 1]
    return corr, pvals


def detrend_sparse(matrix, detectable_bins=None, max_dist=None, max_val=10):
    """pre:256-310 (smooth=False, fun=nanmean) on a sparse matrix: per-diagonal
    mean of the positive pixels between detectable bins, divide, clamp."""
    coo = matrix.tocoo()
    n = coo.shape[0]
    if max_dist is None:
        max_dist = n
    n_diags = min(n, max_dist + 1)
    ok = np.zeros(n, dtype=bool)
    if detectable_bins is None:
        ok[:] = True
    else:
        ok[np.asarray(detectable_bins, dtype=int)] = True
    d = coo.col - coo.row
    use = (d >= 0) & (d < n_diags) & ok[coo.row] & ok[coo.col] & (coo.data > 0)
    tot = np.bincount(d[use], weights=coo.data[use], minlength=n)[:n]
    cnt = np.bincount(d[use], minlength=n)[:n]
    with np.errstate(all="ignore"):
        law = np.where(cnt > 0, tot / np.maximum(cnt, 1), 0.0)
        val = coo.data / law[np.abs(d)]
    if max_val is not None:
        val[val >= max_val] = 1.0
    return sp.csr_matrix((val, (coo.row, coo.col)), shape=coo.shape)
