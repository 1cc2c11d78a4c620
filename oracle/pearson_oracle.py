"""Dense per-window restatement of chromosight's normalised cross-correlation.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity: PINNED against the
reference through tests/golden/*.npz (tests/test_oracle_golden.py).

The reference (chromosight/utils/detection.py) evaluates these formulas with
scipy.sparse Toeplitz products; this file evaluates the same quantities window
by window on dense float64 arrays with scipy.signal.correlate2d, i.e. through
an independent code path.  Every function cites the reference lines it follows.

Notation for a window of N = mk*nk pixels of the (framed) signal S with missing
mask M (True = missing, S is 0 there):
    T(x) = 0 if |x| < 1e-4 else x        (xcorr2's default threshold, det:595,716)
"""
import numpy as np
import scipy.signal as _sig
from scipy.special import erfc as _erfc

XCORR_THRESHOLD = 1e-4  # det:595 (default of xcorr2, used by every call in normxcorr2)
DENOM_EPS = 1e-10       # det:1024, det:1088


def _thr(x, threshold=XCORR_THRESHOLD):
    """det:716 / det:803 - values whose magnitude is below threshold become 0."""
    x = np.array(x, dtype=np.float64, copy=True)
    x[np.abs(x) < threshold] = 0.0
    return x


def truncate_kernel(kernel, prop_info):
    """Rank-truncated kernel equal to U @ V of preprocessing.factorise_kernel
    (pre:810-847): keep the first singular triplets whose cumulated squared
    singular values exceed prop_info of the total."""
    u, s, vt = np.linalg.svd(np.asarray(kernel, dtype=np.float64))
    keep = int(np.flatnonzero(np.cumsum(s ** 2) > prop_info * np.sum(s ** 2))[0]) + 1
    return (u[:, :keep] * s[:keep]) @ vt[:keep, :]


def xcorr2_dense(signal, kernel, threshold=XCORR_THRESHOLD, tsvd=None):
    """det:595-624 + det:726-804: valid-mode cross-correlation, thresholded,
    returned at the signal's shape with zero margins of (k-1)//2."""
    signal = np.asarray(signal, dtype=np.float64)
    kernel = np.asarray(kernel, dtype=np.float64)
    if tsvd is not None:
        kernel = truncate_kernel(kernel, tsvd)
    km, kn = kernel.shape
    sm, sn = signal.shape
    out = np.zeros((sm, sn))
    kh, kw = (km - 1) // 2, (kn - 1) // 2
    valid = _sig.correlate2d(signal, kernel, mode="valid")
    out[kh:kh + valid.shape[0], kw:kw + valid.shape[1]] = valid
    return _thr(out, threshold)


def make_missing_mask_dense(shape, valid_rows, valid_cols, max_dist=None, sym_upper=False):
    """pre:535-633 as a dense boolean array.

    sym_upper: pixel (r, c) is flagged iff 0 <= c - r <= max_dist and bin r or
    bin c is missing (pre:588-627); otherwise whole missing rows and columns
    (pre:629-631)."""
    sm, sn = shape
    miss_r = np.ones(sm, dtype=bool)
    miss_r[np.asarray(valid_rows, dtype=int)] = False
    if sym_upper:
        if sm != sn or len(valid_rows) != len(valid_cols):
            raise ValueError("Rectangular matrices cannot be upper symmetric")
        miss_c = miss_r
        if max_dist is None:
            max_dist = min(shape)
        r = np.arange(sm)[:, None]
        c = np.arange(sn)[None, :]
        band = (c - r >= 0) & (c - r <= max_dist)
        return band & (miss_r[:, None] | miss_c[None, :])
    miss_c = np.ones(sn, dtype=bool)
    miss_c[np.asarray(valid_cols, dtype=int)] = False
    return miss_r[:, None] | miss_c[None, :]


def frame_missing_mask_dense(mask, kernel_shape, sym_upper=False, max_dist=None):
    """pre:404-498 as a dense boolean array of shape
    (ms + 2(mk-1), ns + 2(nk-1))."""
    mask = np.asarray(mask, dtype=bool)
    ms, ns = mask.shape
    mk, nk = kernel_shape
    banded = sym_upper and (max_dist is not None)
    inner = mask.copy()
    if banded:
        # pre:452-454: diag_trim(mask, max_dist + max(nk, mk)) keeps diagonals 0..n
        r = np.arange(ms)[:, None]
        c = np.arange(ns)[None, :]
        d = c - r
        inner &= (d >= 0) & (d <= max_dist + max(nk, mk))
        max_m, max_n = max_dist + mk, max_dist + nk
    H, W = ms + 2 * (mk - 1), ns + 2 * (nk - 1)
    framed = np.zeros((H, W), dtype=bool)
    framed[mk - 1:mk - 1 + ms, nk - 1:nk - 1 + ns] = inner
    if banded:
        # pre:461-463 top margin up to the scan distance
        framed[: mk - 1, nk - 1: nk - 1 + min(max_n, ns)] = True
        # pre:475 right margin, last max_m + 1 framed rows
        if nk > 1:
            framed[max(0, H - (max_m + 1)):, W - (nk - 1):] = True
        # pre:477 top-left corner
        framed[: mk - 1, : nk - 1] = True
    else:
        framed[: mk - 1, :] = True
        framed[H - (mk - 1):, :] = True
        framed[:, : nk - 1] = True
        framed[:, W - (nk - 1):] = True
    if sym_upper:
        # pre:483-497: big_k diagonals below the main diagonal of the framed matrix
        big_k = max(nk, mk)
        Y = np.arange(H)[:, None]
        X = np.arange(W)[None, :]
        framed |= (Y - X >= 1) & (Y - X <= big_k)
    return framed


def corr_to_log10_pval(corr, n_obs):
    """stats:43-81: two-sided log10 p-value of a Pearson coefficient through
    Fisher's z.  2*Phi(-|z|) == erfc(|z|/sqrt(2)); scipy's ndtr returns exactly 0
    once (|z|/sqrt 2)^2 exceeds log(DBL_MAX) (cephes underflow rule)."""
    corr = np.asarray(corr, dtype=np.float64)
    n_obs = np.asarray(n_obs, dtype=np.float64)
    with np.errstate(all="ignore"):
        z = np.abs(np.arctanh(corr) * np.sqrt(n_obs - 3.0))
        a = z * np.sqrt(0.5)
        p = _erfc(a)
        p = np.where(a * a > 7.09782712893383996843e2, 0.0, p)
        p = np.where(np.isnan(z), np.nan, p)
        return np.log10(p)


def normxcorr2_dense(
    signal,
    kernel,
    max_dist=None,
    sym_upper=False,
    full=False,
    missing_mask=None,
    missing_tol=0.75,
    tsvd=None,
    pval=False,
    return_nobs=False,
    return_cond=False,
):
    """det:917-1131 (`_normxcorr2_sparse`) restated on dense float64 arrays.

    signal : (ms, ns) array; missing_mask : (ms, ns) bool array or None.
    Returns (corr, log10_pvals) as dense (ms, ns) arrays; pixels the reference
    leaves unstored are 0 in both."""
    S = np.asarray(signal, dtype=np.float64)
    K = np.asarray(kernel, dtype=np.float64)
    mk, nk = K.shape
    ms, ns = S.shape
    N = mk * nk
    ones = np.ones((mk, nk))
    if full:
        F = np.zeros((ms + 2 * (mk - 1), ns + 2 * (nk - 1)))
        F[mk - 1:mk - 1 + ms, nk - 1:nk - 1 + ns] = S          # det:979-985
        M = None
        if missing_mask is not None:
            M = frame_missing_mask_dense(missing_mask, K.shape, sym_upper, max_dist)  # det:989
    else:
        F = S
        M = None if missing_mask is None else np.asarray(missing_mask, dtype=bool)

    Kc = K if tsvd is None else truncate_kernel(K, tsvd)
    with np.errstate(all="ignore"):
        mean_s = xcorr2_dense(F, ones / N)                      # det:1005 / det:1050
        mean_s2 = xcorr2_dense(F ** 2, ones / N)                # det:1007 / det:1055
        mean_sk = xcorr2_dense(F, Kc / N)                       # det:1018 / det:1082
        if M is None:
            k_mean, k_std = float(K.mean()), float(K.std())
            denom = np.sqrt(mean_s2 - mean_s ** 2) * k_std      # det:1007-1008
            num = mean_sk - mean_s * k_mean                     # det:1018
            bad = ~(np.abs(denom) >= DENOM_EPS)                 # det:1011 (NaN -> dropped later)
            r = np.where(bad, 0.0, num / denom)
            n_obs = np.full(F.shape, float(N))
            cond = np.sqrt(mean_s2 / (mean_s2 - mean_s ** 2))
        else:
            k_sum, k2_sum = K.sum(), (K ** 2).sum()
            k_mean, k2_mean = k_sum / N, k2_sum / N
            Mf = M.astype(np.float64)
            K2c = K ** 2 if tsvd is None else truncate_kernel(K ** 2, tsvd)
            n_miss = xcorr2_dense(Mf, ones)                     # det:1030
            has = n_miss != 0
            n_pres = N - n_miss                                 # det:1033
            km_wm = (k_sum - xcorr2_dense(Mf, Kc)) / n_pres     # det:1035-1040
            k2m_wm = (k2_sum - xcorr2_dense(Mf, K2c)) / n_pres  # det:1041-1046
            scale = np.where(has, N / n_pres, 1.0)
            m_s = mean_s * scale                                # det:1051-1053
            m_s2 = mean_s2 * scale                              # det:1056-1058
            var_k = np.where(has, k2m_wm - km_wm ** 2, k2_mean - k_mean ** 2)
            denom = np.sqrt((m_s2 - m_s ** 2) * var_k)          # det:1060-1066
            few = has & (n_pres < int((1 - missing_tol) * N))   # det:1069-1072
            denom = np.where(few, 0.0, denom)
            # det:1075-1085.  The reference rescales the masked entries through
            # 1/(kernel_mean*kernel_size): a zero-mean kernel turns them into NaN.
            cov_plain = mean_sk - mean_s * k_mean
            cov_miss = (mean_sk - mean_s * km_wm) * scale
            if k_mean == 0:
                cov_miss = np.full(F.shape, np.nan)
            num = np.where(has, cov_miss, cov_plain)
            bad = ~(np.abs(denom) >= DENOM_EPS)                 # det:1088-1091
            r = np.where(bad, 0.0, num / denom)
            cond = np.sqrt(m_s2 / (m_s2 - m_s ** 2))
            if full:
                n_obs = np.where(has & (n_pres != 0), n_pres, float(N))  # det:1110-1116
            else:
                n_obs = np.full(F.shape, float(N))                      # det:1120-1121
        if sym_upper:
            r = np.triu(r)                                      # det:1098-1099
        r[~np.isfinite(r)] = 0.0                                # det:1101
        r = np.clip(r, -1.0, 1.0)                               # det:1105-1106
        if pval:
            p = corr_to_log10_pval(r, n_obs)                    # det:1108-1121
            p = np.where(r == 0.0, 0.0, p)                      # unstored pixels
        else:
            p = None
    if full:
        r = r[mk - 1:mk - 1 + ms, nk - 1:nk - 1 + ns]           # det:1124-1129
        if p is not None:
            p = p[mk - 1:mk - 1 + ms, nk - 1:nk - 1 + ns]
        n_obs = n_obs[mk - 1:mk - 1 + ms, nk - 1:nk - 1 + ns]
        cond = cond[mk - 1:mk - 1 + ms, nk - 1:nk - 1 + ns]
    if return_cond:
        # rms / std of the present pixels of each window: how much a relative
        # perturbation of the signal (e.g. float32 storage) is amplified in r
        return r, p, n_obs, np.where(np.isfinite(cond), cond, np.inf)
    if return_nobs:
        return r, p, n_obs
    return r, p


def distance_law_dense(matrix, detectable_bins=None, max_dist=None):
    """pre:129-197 with smooth=False, fun=nanmean: mean, per upper diagonal
    d <= max_dist, of the strictly positive pixels whose two bins are detectable."""
    A = np.asarray(matrix, dtype=np.float64)
    n = A.shape[0]
    if max_dist is None:
        max_dist = n
    n_diags = min(n, max_dist + 1)
    ok = np.zeros(n, dtype=bool)
    if detectable_bins is None:
        ok[:] = True
    else:
        ok[np.asarray(detectable_bins, dtype=int)] = True
    dist = np.zeros(n)
    for d in range(n_diags):
        vals = np.diagonal(A, d)[ok[: n - d] & ok[d:]]
        vals = vals[vals > 0]
        dist[d] = vals.mean() if len(vals) else np.nan
    return dist


def detrend_dense(matrix, detectable_bins=None, max_dist=None, max_val=10):
    """pre:256-310 with smooth=False: divide every stored pixel by the distance
    law at |row - col|; results >= max_val (including inf from a zero law)
    become 1.  Zero pixels stay zero (they are not stored in the reference)."""
    A = np.asarray(matrix, dtype=np.float64)
    y = distance_law_dense(A, detectable_bins, max_dist)
    y[np.isnan(y)] = 0.0
    r = np.arange(A.shape[0])[:, None]
    c = np.arange(A.shape[1])[None, :]
    with np.errstate(all="ignore"):
        out = np.where(A != 0, A / y[np.abs(r - c)], 0.0)
    if max_val is not None:
        out[out >= max_val] = 1.0
    return out
