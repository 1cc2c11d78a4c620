"""Mirror of the parts of chromosight.utils.preprocessing that sit on the hot
path: same names, same arguments, same error
behaviour.  The O(nnz) arithmetic of the distance law and of detrending runs in
CUDA (csrc/detrend.cu); masks, trimming and padding are host bookkeeping on
scipy.sparse objects, as in the reference.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from .. import _cuda, _lib


# --------------------------------------------------------------------------- host helpers
def valid_to_missing(valid, size):
    """Complement of an index list within range(size) (pre:850-875)."""
    flags = np.ones(size, dtype=bool)
    try:
        flags[valid] = False
    except IndexError:
        pass
    return np.flatnonzero(flags)


def diag_trim(mat, n):
    """Keep diagonals 0..n (inclusive) of the upper triangle (pre:93-126)."""
    if sp.issparse(mat):
        if mat.format != "csr":
            raise ValueError("input type must be scipy.sparse.csr_matrix")
        coo = mat.tocoo()
        d = coo.col - coo.row
        keep = (d >= 0) & (d <= n)
        return sp.csr_matrix((coo.data[keep], (coo.row[keep], coo.col[keep])), shape=mat.shape)
    out = np.array(mat, copy=True)
    rows = np.arange(out.shape[0])[:, None]
    cols = np.arange(out.shape[1])[None, :]
    # the dense variant of the reference blanks diagonals n.. of the upper part
    # only and leaves the lower triangle untouched (pre:119-124)
    out[(cols - rows) >= n] = 0
    return out


def zero_pad_sparse(mat, margin_h, margin_v, fmt="coo"):
    """Zero margins around a sparse matrix (pre:636-676)."""
    coo = sp.coo_matrix(mat)
    sm, sn = coo.shape
    out = sp.coo_matrix(
        (coo.data, (coo.row + margin_v, coo.col + margin_h)),
        shape=(sm + 2 * margin_v, sn + 2 * margin_h),
    )
    # the reference always hands back CSR, whatever fmt says (pre:671-676)
    return out.tocsr()


def missing_geometry(shape, valid_rows, valid_cols, max_dist=None, sym_upper=False):
    """The ingredients of make_missing_mask (pre:535-633) without the matrix: boolean
    vectors of missing rows / columns and the diagonals col - row on which a missing bin
    flags its pixels (None = unbounded).  This is what the device needs of the mask."""
    sm, sn = shape
    if sym_upper and (sm != sn or len(valid_rows) != len(valid_cols)):
        raise ValueError("Rectangular matrices cannot be upper symmetric")
    fr = np.ones(sm, dtype=bool)
    fr[np.asarray(valid_rows, dtype=np.int64)] = False
    if sym_upper:
        if max_dist is None:
            max_dist = min(shape)
        return fr, fr, 0, int(max_dist)
    fc = np.ones(sn, dtype=bool)
    fc[np.asarray(valid_cols, dtype=np.int64)] = False
    return fr, fc, None, None


def make_missing_mask(shape, valid_rows, valid_cols, max_dist=None, sym_upper=False):
    """Sparse boolean mask of the pixels that belong to missing bins (pre:535-633).

    The returned matrix remembers how it was made (`_cs_geometry`): normxcorr2 then hands
    the device two bit vectors instead of the pixel list."""
    sm, sn = shape
    if sym_upper and (sm != sn or len(valid_rows) != len(valid_cols)):
        raise ValueError("Rectangular matrices cannot be upper symmetric")
    miss_r = valid_to_missing(valid_rows, sm)
    geometry = missing_geometry(shape, valid_rows, valid_cols, max_dist, sym_upper)
    if sym_upper:
        if max_dist is None:
            max_dist = min(shape)
        shifts = np.arange(max_dist + 1)
        # pixels above each missing bin (same column) and to its right (same row)
        rows = np.concatenate([(miss_r[:, None] - shifts[None, :]).ravel(),
                               np.repeat(miss_r, max_dist + 1)])
        cols = np.concatenate([np.repeat(miss_r, max_dist + 1),
                               (miss_r[:, None] + shifts[None, :]).ravel()])
        ok = (rows >= 0) & (rows < sm) & (cols >= 0) & (cols < sm)
        mask = sp.coo_matrix((np.ones(ok.sum(), dtype=bool), (rows[ok], cols[ok])),
                             shape=shape, dtype=bool).tocsr()
        mask._cs_geometry = geometry + (mask.nnz,)
        return mask
    miss_c = valid_to_missing(valid_cols, sn)
    fr = np.zeros(sm, dtype=bool)
    fr[miss_r] = True
    # whole rows, then the remaining pixels of whole columns
    r1 = np.repeat(miss_r, sn)
    c1 = np.tile(np.arange(sn), len(miss_r))
    good_rows = np.flatnonzero(~fr)
    r2 = np.tile(good_rows, len(miss_c))
    c2 = np.repeat(miss_c, len(good_rows))
    rows = np.concatenate([r1, r2])
    cols = np.concatenate([c1, c2])
    mask = sp.coo_matrix((np.ones(len(rows), dtype=bool), (rows, cols)), shape=shape,
                         dtype=bool).tocsr()
    mask._cs_geometry = geometry + (mask.nnz,)
    return mask


def mask_geometry(mask, sym_upper=False):
    """(miss_rows, miss_cols, dlo, dhi) if `mask` (scipy sparse, bool) is exactly a mask
    make_missing_mask can build, else None.  Masks made by this module carry the answer;
    foreign ones (e.g. the reference's own make_missing_mask) are recognised from their
    pattern in O(nnz)."""
    tag = getattr(mask, "_cs_geometry", None)
    if tag is not None and tag[4] == mask.nnz and len(tag[0]) == mask.shape[0] \
            and len(tag[1]) == mask.shape[1]:
        return tag[:4]
    csr = mask.tocsr()
    if not csr.has_canonical_format:
        csr = csr.copy()
        csr.sum_duplicates()
    if csr.nnz and not np.all(csr.data):
        csr = csr.copy()
        csr.eliminate_zeros()
    sm, sn = csr.shape
    counts = np.diff(csr.indptr)
    rows = np.repeat(np.arange(sm), counts)
    cols = csr.indices
    if csr.nnz == 0:
        return np.zeros(sm, bool), np.zeros(sn, bool), (0 if sym_upper else None), (0 if sym_upper else None)
    d = cols.astype(np.int64) - rows
    if sm == sn and d.min() >= 0:
        # banded upper-triangular generator: bin b is missing iff pixel (b, b) is flagged
        miss = np.zeros(sm, dtype=bool)
        miss[rows[d == 0]] = True
        md = int(d.max())
        cum = np.concatenate([[0], np.cumsum(miss)])
        r = np.arange(sm)
        last = np.minimum(r + md, sm - 1)
        expect = np.where(miss, last - r + 1, cum[last + 1] - cum[r])
        if np.array_equal(expect, counts) and np.all(miss[rows] | miss[cols]):
            return miss, miss, 0, md
    # whole rows and whole columns
    fr = counts == sn
    fc = np.bincount(cols, minlength=sn) == sm
    n_r, n_c = int(fr.sum()), int(fc.sum())
    if csr.nnz == n_r * sn + n_c * (sm - n_r) and np.all(fr[rows] | fc[cols]):
        return fr, fc, None, None
    return None


def frame_missing_mask(mask, kernel_shape, sym_upper=False, max_dist=None):
    """Missing mask of the framed signal (pre:404-498).  The CUDA path builds this
    frame itself (csrc/image.cu); this host version exists for API parity."""
    if mask.dtype != bool:
        raise ValueError("Mask must contain boolean values")
    if not sp.issparse(mask):
        raise ValueError("Mask must be a sparse matrix")
    ms, ns = mask.shape
    mk, nk = kernel_shape
    banded = sym_upper and (max_dist is not None)
    coo = mask.tocoo()
    keep = coo.data != 0
    r, c = coo.row[keep], coo.col[keep]
    if banded:
        d = c - r
        ok = (d >= 0) & (d <= max_dist + max(nk, mk))
        r, c = r[ok], c[ok]
    H, W = ms + 2 * (mk - 1), ns + 2 * (nk - 1)
    framed = sp.coo_matrix((np.ones(len(r), dtype=bool), (r + mk - 1, c + nk - 1)),
                           shape=(H, W), dtype=bool).tolil()
    if banded:
        max_m, max_n = max_dist + mk, max_dist + nk
        framed[: mk - 1, nk - 1: nk - 1 + min(max_n, ns)] = True
        if nk > 1:
            framed[max(0, H - (max_m + 1)):, W - (nk - 1):] = True
        framed[: mk - 1, : nk - 1] = True
    else:
        framed[: mk - 1, :] = True
        framed[H - (mk - 1):, :] = True
        framed[:, : nk - 1] = True
        framed[:, W - (nk - 1):] = True
    framed = framed.tocsr()
    if sym_upper:
        big_k = max(nk, mk)
        framed = framed + sp.diags(np.ones(big_k), -np.arange(1, big_k + 1), shape=(H, W),
                                   format="csr", dtype=bool)
        framed = framed.astype(bool)
    return framed.tocsr()


def check_missing_mask(signal, mask):
    """Raise ValueError when the signal is non-zero under the mask (pre:501-532)."""
    if sp.issparse(mask):
        m = mask.tocoo()
        r, c = m.row[m.data != 0], m.col[m.data != 0]
        vals = np.asarray(signal[r, c]).ravel() if len(r) else np.zeros(0)
        n_bad = int(np.count_nonzero(np.abs(vals) > 0))
        if n_bad:
            raise ValueError("There are", n_bad, "non-zero elements reported as missing.")
    else:
        tot = np.sum(np.abs(np.asarray(signal)[np.asarray(mask) > 0]))
        if tot > 1e-10:
            raise ValueError("There are", str(tot), "non-zero elements reported as missing.")


def factorise_kernel(kernel, prop_info=0.999):
    """Truncated SVD factors (U sqrt(s), sqrt(s) V) keeping `prop_info` of the
    squared singular values (pre:810-847)."""
    u, s, vt = np.linalg.svd(np.asarray(kernel, dtype=float))
    keep = int(np.flatnonzero(np.cumsum(s ** 2) > prop_info * np.sum(s ** 2))[0]) + 1
    root = np.sqrt(s[:keep])
    return u[:, :keep] * root[None, :], vt[:keep, :] * root[:, None]


def truncate_kernel(kernel, prop_info):
    """U @ V of factorise_kernel: the dense kernel the factorised convolution is
    equivalent to (det:648-665)."""
    left, right = factorise_kernel(kernel, prop_info)
    return left @ right


def resize_kernel(kernel, kernel_res=None, signal_res=None, factor=None, min_size=7, quiet=False):
    """Rescale a square, odd kernel by `factor` (or kernel_res / signal_res) with linear
    interpolation; the result is kept odd and at least min_size wide (pre:731-808)."""
    import scipy.ndimage as ndi
    kernel = np.asarray(kernel, dtype=float)
    km, kn = kernel.shape
    if km != kn:
        raise ValueError("kernel must be square.")
    if km % 2 == 0:
        raise ValueError("kernel size must be odd.")
    if factor is not None:
        if kernel_res is not None or signal_res is not None:
            raise ValueError("factor is mutually exclusive with resolution "
                             "parameters (kernel_res and signal_res).")
        scale = factor
    else:
        if kernel_res is None or signal_res is None:
            raise ValueError("You must provide either a resize factor or the signal and "
                             "kernel resolutions.")
        scale = kernel_res / signal_res
    scale = max(scale, min_size / km)
    out = ndi.zoom(kernel, scale, order=1)
    if out.shape[0] % 2 == 0:
        # one pixel smaller keeps the centre on a pixel (pre:796-806)
        out = ndi.zoom(kernel, (out.shape[0] - 1) / km, order=1)
    return out


def crop_kernel(kernel, target_size):
    """Trim equal margins so that the kernel is no larger than target_size (made odd by
    rounding up), pre:679-728."""
    tm, tn = (d + 1 - d % 2 for d in target_size)
    sm, sn = kernel.shape
    mr = (sm - tm) // 2 if sm > tm else 0
    mc = (sn - tn) // 2 if sn > tn else 0
    return kernel[mr:sm - mr, mc:sn - mc]


def ztransform(matrix):
    """Global z-score of the stored values (pre:313-334)."""
    out = matrix.copy()
    out.data = (out.data - np.mean(out.data)) / np.std(out.data)
    return out


# --------------------------------------------------------------------------- CUDA path
def divide_by_median(data):
    """Inter-chromosomal normalisation (cm:598-601): NaN -> 0, then every stored value divided by
    the median of the stored values (np.nanmedian semantics: mean of the two middle values),
    sorted and divided on the device."""
    t = _cuda.require_cuda()
    d = _cuda.to_device(np.asarray(data, dtype=np.float64))
    d = t.nan_to_num(d, nan=0.0)
    n = d.numel()
    if n == 0:
        return np.asarray(data, dtype=np.float64)
    srt = t.sort(d).values
    med = (srt[(n - 1) // 2] + srt[n // 2]) / 2.0
    return (d / med).cpu().numpy()


def detrend_band_device(indptr, indices, data, n, detectable_bins=None, max_dist=None, max_val=10):
    """`detrend` (pre:256-310) of an upper-band CSR given as host arrays, result left in HBM:
    -> _cuda.DeviceCSR with the detrended values (NaN -> 0, cm:537-538).  The arrays go to the
    device once; law, division and clamp run there (csrc/detrend.cu); nothing comes back.
    This is ContactMap.create_mat's intra branch (cm:607-624) fused with the upload of
    pattern_detector: diag_trim is implicit (the band holds diagonals 0..max_dist only)."""
    t = _cuda.require_cuda()
    lib = _lib.load()
    n = int(n)
    d_indptr = _cuda.to_device(indptr, np.int64)
    d_indices = _cuda.to_device(indices, np.int32)
    d_data = _cuda.to_device(data, np.float64)
    nnz = int(indptr[-1])

    class _Shape:
        shape = (n, n)
    d_law, _, _ = _law_device(_Shape, (d_indptr, d_indices, d_data), detectable_bins, max_dist)
    out = _cuda.empty(nnz, t.float64)
    if nnz:
        _lib.check(lib.cs_detrend_apply(_cuda.ptr(d_indptr), _cuda.ptr(d_indices), _cuda.ptr(d_data),
                                        _cuda.ptr(out), n, _cuda.ptr(d_law), n,
                                        C.c_double(-1.0 if max_val is None else float(max_val)),
                                        _cuda.stream_ptr()))
        out = t.where(t.isnan(out), t.zeros((), dtype=out.dtype, device=out.device), out)
    dmax = n - 1 if max_dist is None else int(min(max_dist, n - 1))
    return _cuda.DeviceCSR((n, n), indptr, d_indices, out, (0, dmax))


def divide_by_median_device(indptr, indices, data, shape):
    """Inter-chromosomal normalisation (cm:598-601) of a CSR block given as host arrays, result
    left in HBM as a _cuda.DeviceCSR (see divide_by_median)."""
    t = _cuda.require_cuda()
    d_indices = _cuda.to_device(indices, np.int32)
    d = _cuda.to_device(np.asarray(data, dtype=np.float64))
    d = t.where(t.isnan(d), t.zeros((), dtype=d.dtype, device=d.device), d)
    n = d.numel()
    if n:
        srt = t.sort(d).values
        d = d / ((srt[(n - 1) // 2] + srt[n // 2]) / 2.0)
    return _cuda.DeviceCSR(shape, indptr, d_indices, d, (-(shape[0] - 1), shape[1] - 1))


def _csr_device(csr):
    indptr = _cuda.to_device(csr.indptr, np.int64)
    indices = _cuda.to_device(csr.indices, np.int32)
    data = _cuda.to_device(csr.data, np.float64)
    return indptr, indices, data


def _law_device(csr, d_csr, detectable_bins, max_dist):
    """Distance law of a CSR matrix already on the device -> (torch tensor of
    length n, float64)."""
    t = _cuda.require_cuda()
    lib = _lib.load()
    n = csr.shape[0]
    if max_dist is None:
        max_dist = n
    n_diags = int(min(n, max_dist + 1))
    if detectable_bins is None:
        d_detect = None
    else:
        flags = np.zeros(n, dtype=np.uint8)
        flags[np.asarray(detectable_bins)] = 1
        d_detect = _cuda.to_device(flags)
    d_sum = _cuda.empty(n_diags, t.float64)
    d_cnt = _cuda.empty(n_diags, t.int64)
    d_law = _cuda.empty(n, t.float64)
    indptr, indices, data = d_csr
    _lib.check(lib.cs_distance_law(_cuda.ptr(indptr), _cuda.ptr(indices), _cuda.ptr(data), n,
                                   _cuda.ptr(d_detect), n_diags, _cuda.ptr(d_sum),
                                   _cuda.ptr(d_cnt), _cuda.ptr(d_law), _cuda.stream_ptr()))
    return d_law, d_cnt, n_diags


def _law_host_fun(csr, detectable_bins, max_dist, fun):
    """Distance law with an arbitrary reduction `fun` (pre:129-136 lets the caller pass any
    callable): the qualifying pixels of each diagonal (upper, <= max_dist, strictly positive,
    both bins detectable; pre:178-188) are grouped on the host and reduced by `fun` itself --
    an arbitrary Python callable cannot run on the device.  The default (np.nanmean) never
    comes here."""
    n = csr.shape[0]
    if max_dist is None:
        max_dist = n
    n_diags = int(min(n, max_dist + 1))
    coo = csr.tocoo()
    d = coo.col.astype(np.int64) - coo.row
    ok = (d >= 0) & (d < n_diags) & (coo.data > 0)
    if detectable_bins is not None:
        good = np.zeros(n, dtype=bool)
        good[np.asarray(detectable_bins)] = True
        ok &= good[coo.row] & good[coo.col]
    d, v = d[ok], coo.data[ok]
    order = np.argsort(d, kind="stable")
    d, v = d[order], v[order]
    starts = np.searchsorted(d, np.arange(n_diags + 1))
    law = np.zeros(n)
    for k in range(n_diags):
        seg = v[starts[k]:starts[k + 1]]
        law[k] = fun(seg) if len(seg) else np.nan
    return law, n_diags


def distance_law(matrix, detectable_bins=None, max_dist=None, smooth=True, fun=np.nanmean):
    """Average contact value per upper diagonal (pre:129-197)."""
    csr = sp.csr_matrix(matrix)
    if not csr.has_canonical_format:
        csr = csr.copy()
        csr.sum_duplicates()
    n = csr.shape[0]
    if fun is not np.nanmean:
        law, n_diags = _law_host_fun(csr, detectable_bins, max_dist, fun)
    else:
        d_law, d_cnt, n_diags = _law_device(csr, _csr_device(csr), detectable_bins, max_dist)
        law = d_law.cpu().numpy()
        cnt = d_cnt.cpu().numpy()
        # the reference reports NaN for diagonals without any usable pixel (nanmean of
        # an empty slice); the device array already holds the 0 that detrend needs
        law[:n_diags][cnt == 0] = np.nan
    if smooth and n > 2:
        from sklearn.isotonic import IsotonicRegression
        law[~np.isfinite(law)] = 0
        law = IsotonicRegression(increasing=False).fit_transform(range(len(law)), law)
    return law


def detrend(matrix, detectable_bins=None, max_dist=None, smooth=False, fun=np.nanmean, max_val=10):
    """Divide each pixel by the distance law of its diagonal (pre:256-310)."""
    t = _cuda.require_cuda()
    lib = _lib.load()
    csr = sp.csr_matrix(matrix, dtype=np.float64)
    if not csr.has_canonical_format:
        csr = csr.copy()
        csr.sum_duplicates()
    n = csr.shape[0]
    d_csr = _csr_device(csr)
    if smooth or fun is not np.nanmean:
        # isotonic smoothing (sklearn) and custom reductions are host code, as in the reference;
        # the division by the law still runs on the device
        law = distance_law(csr, detectable_bins, max_dist, smooth=smooth, fun=fun)
        law[np.isnan(law)] = 0.0
        d_law = _cuda.to_device(law, np.float64)
    else:
        d_law, _, _ = _law_device(csr, d_csr, detectable_bins, max_dist)
    out = _cuda.empty(csr.nnz, t.float64)
    if csr.nnz:
        _lib.check(lib.cs_detrend_apply(_cuda.ptr(d_csr[0]), _cuda.ptr(d_csr[1]), _cuda.ptr(d_csr[2]),
                                        _cuda.ptr(out), n, _cuda.ptr(d_law), n,
                                        C.c_double(-1.0 if max_val is None else float(max_val)),
                                        _cuda.stream_ptr()))
    return sp.csr_matrix((out.cpu().numpy(), csr.indices.copy(), csr.indptr.copy()), shape=csr.shape)
