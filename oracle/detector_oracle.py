"""Dense restatement of pattern_detector's post-processing (det:177-345) and of
validate_patterns (det:18-155).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity: PINNED against the
unmodified reference through tests/golden/detector_*.npz
(tests/test_oracle_golden.py::test_detector_oracle_matches_reference).
"""
import numpy as np

from . import pearson_oracle as po


def validate_patterns_dense(coords, mat, conv, missing_rows, missing_cols, kernel_shape, zero_tol,
                            missing_tol):
    """det:18-155 on dense arrays.  `mat` is the (padded, NaN sub-diagonal) matrix the
    windows are cut from, `conv` the (padded) correlation map, missing_* boolean flags per
    padded row / column.  Returns (windows, valid, score)."""
    win_h, win_w = kernel_shape
    half_h, half_w = win_h // 2 + 1, win_w // 2 + 1
    P = len(coords)
    windows = np.full((P, win_h, win_w), np.nan)
    valid = np.zeros(P, dtype=bool)
    score = np.full(P, np.nan)
    for i, (p1, p2) in enumerate(np.asarray(coords, dtype=int)):
        high, low = p1 - half_h + 1, p1 + half_h
        left, right = p2 - half_w + 1, p2 + half_w
        if not (high >= 0 and low < mat.shape[0] and left >= 0 and right < mat.shape[1]):  # det:97-102
            continue
        w = mat[high:low, left:right].copy()
        w[missing_rows[high:low], :] = np.nan             # det:117-118
        w[:, missing_cols[left:right]] = np.nan
        tot = w.size
        n_zero = np.sum(w == 0)
        n_miss = np.sum(~np.isfinite(w))
        with np.errstate(all="ignore"):
            prop_undetected = n_miss / tot
            prop_zero = np.float64(n_zero) / np.float64(tot - n_miss)
        if prop_undetected < missing_tol and prop_zero < zero_tol:   # det:133
            valid[i] = True
            windows[i] = w
            score[i] = conv[p1, p2]
    return windows, valid, score


def pattern_detector_dense(matrix, detectable_bins, max_dist, inter, config, kernel, pick_foci,
                           coords=None, full=False):
    """det:177-345 on dense arrays; `pick_foci` is the foci picker to use (a callable with
    the reference's signature).  Returns (bin1, bin2, score, pvalue, windows) or None."""
    A = np.asarray(matrix, dtype=np.float64)
    K = np.asarray(kernel, dtype=np.float64)
    km, kn = K.shape
    kh, kw = (km - 1) // 2, (kn - 1) // 2
    quantify = coords is not None
    if min(A.shape) <= max(K.shape):
        return None
    mask = None
    if full:
        mask = po.make_missing_mask_dense(A.shape, detectable_bins[0], detectable_bins[1],
                                          max_dist=max_dist, sym_upper=not inter)
    r, logp = po.normxcorr2_dense(A, K, max_dist=max_dist, sym_upper=not inter, full=full,
                                  missing_mask=mask, pval=True,
                                  missing_tol=config["max_perc_undetected"] / 100)
    conv = r.copy()
    if not inter:                                         # det:270: diagonals 0..max_dist
        i, j = np.indices(conv.shape)
        conv[(j - i < 0) | (j - i > max_dist)] = 0
    if not quantify:
        import scipy.sparse as sp
        coords, _ = pick_foci(sp.coo_matrix(conv), config["pearson"])
        if coords is None:
            return None
    coords = np.array(coords, dtype=int).reshape(-1, 2)
    mat = A.copy()
    det_r = np.asarray(detectable_bins[0]).copy()
    det_c = np.asarray(detectable_bins[1]).copy()
    if full:                                              # det:291-298 (margins as the reference passes them)
        mat = np.pad(mat, ((kw, kw), (kh, kh)))
        conv = np.pad(conv, ((kw, kw), (kh, kh)))
        det_r = det_r + kh
        det_c = det_c + kw
        coords = coords + np.array([kh, kw])
    if not inter:                                         # det:300-315
        big_k = max(km, kn)
        i, j = np.indices(mat.shape)
        mat[(i - j >= 1) & (i - j <= big_k)] = np.nan
        if config["max_dist"] == 0:
            coords[:, 0] = coords[:, 1]
    miss_r = np.ones(mat.shape[0], dtype=bool)
    miss_r[det_r[det_r < mat.shape[0]]] = False
    miss_c = np.ones(mat.shape[1], dtype=bool)
    miss_c[det_c[det_c < mat.shape[1]]] = False
    windows, valid, score = validate_patterns_dense(
        coords, mat, conv, miss_r, miss_c, K.shape, config["max_perc_zero"] / 100,
        config["max_perc_undetected"] / 100)
    if full:
        coords = coords - np.array([kh, kw])
    if not quantify:
        coords, windows, score = coords[valid], windows[valid], score[valid]
    with np.errstate(all="ignore"):
        pvalue = 10.0 ** logp[coords[:, 0], coords[:, 1]]
    return coords[:, 0], coords[:, 1], score, pvalue, windows
