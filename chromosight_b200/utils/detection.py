"""Drop-in mirror of chromosight.utils.detection for the hot path.

`normxcorr2` and `xcorr2` keep the reference's signatures, return types and
error behaviour (det:595-624, det:807-914) but run on a B200 through
libchromosight_b200.so: the CSR signal is densified into a float32 band in HBM,
one fused CUDA kernel evaluates every window, and the non-zero scores come
back as the same scipy CSR objects the reference returns.  There is no CPU
implementation behind these functions.

The callers of the hot path (pick_foci & co, det:387-592) are host code, as in
the reference.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from .. import _lib
from . import preprocessing as preproc

# statistics of the last hot-path call (device-side timings, windows evaluated)
last_call_stats = {}
# "auto": masks that make_missing_mask can build reach the device as two bit vectors, any other
# as its pixel list; "pixels" forces the pixel list (tests compare the two device paths)
MASK_FORM = "auto"


def _device_index():
    from .. import _cuda
    t = _cuda.require_cuda()
    return int(t.cuda.current_device())


def _as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _canonical_csr(mat, dtype):
    csr = mat.tocsr() if sp.issparse(mat) else sp.csr_matrix(mat)
    if csr.dtype != dtype:
        csr = csr.astype(dtype)
    if not csr.has_canonical_format:
        csr = csr.copy()
        csr.sum_duplicates()
    return csr


def _kernel_desc(kernel, tsvd, keep):
    """Fill a cs_kernel_desc; `keep` collects the arrays that must outlive the call."""
    K = _as_f64(kernel)
    if tsvd is not None:
        k_corr = _as_f64(preproc.truncate_kernel(K, tsvd))
        k2_mask = _as_f64(preproc.truncate_kernel(K ** 2, tsvd))
    else:
        k_corr = K
        k2_mask = _as_f64(K ** 2)
    keep.extend([K, k_corr, k2_mask])
    d = _lib.KernelDesc()
    d.kh, d.kw = K.shape
    d.k_corr = k_corr.ctypes.data
    d.k_mask = k_corr.ctypes.data
    d.k2_mask = k2_mask.ctypes.data
    d.k_sum = float(K.sum())
    d.k2_sum = float((K ** 2).sum())
    d.k_mean = float(K.mean())
    d.k_std = float(K.std())
    return d


def _build_args(csr, kernel, mask_csr, max_dist, sym_upper, full, missing_tol, tsvd, pval,
                raw_xcorr=False, threshold=1e-4, trim_to_max_dist=False, device=None,
                geometry=None, out_rows=None):
    """cs_normxcorr2_args of one call; the second value keeps the host arrays alive.
    The missing mask is either a pixel mask (`mask_csr`) or the ingredients of
    make_missing_mask (`geometry` = (miss_rows, miss_cols, dlo, dhi), pre:535-633)."""
    keep = []
    a = _lib.Normxcorr2Args()
    a.rows, a.cols = csr.shape
    on_device = hasattr(csr, "d_indices")       # _cuda.DeviceCSR: entries already in HBM
    if on_device:
        indptr = csr.indptr
        keep.extend([indptr, csr.d_indices, csr.d_data])
        a.indptr, a.indices, a.data = indptr.ctypes.data, csr.d_indices.data_ptr(), csr.d_data.data_ptr()
        a.device_payload = 1
    else:
        indptr = np.ascontiguousarray(csr.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(csr.indices, dtype=np.int32)
        data = _as_f64(csr.data)
        keep.extend([indptr, indices, data])
        a.indptr, a.indices, a.data = indptr.ctypes.data, indices.ctypes.data, data.ctypes.data
    a.has_mask = 0
    if geometry is not None:
        miss_r = np.ascontiguousarray(geometry[0], dtype=np.uint8)
        miss_c = np.ascontiguousarray(geometry[1], dtype=np.uint8)
        if miss_r.shape != (csr.shape[0],) or miss_c.shape != (csr.shape[1],):
            raise ValueError("Signal and missing mask do not have the same shape")
        keep.extend([miss_r, miss_c])
        a.has_mask = 2
        a.miss_row, a.miss_col = miss_r.ctypes.data, miss_c.ctypes.data
        a.mask_dlo = -(2 ** 31) if geometry[2] is None else int(geometry[2])
        a.mask_dhi = 2 ** 31 - 1 if geometry[3] is None else int(geometry[3])
    elif mask_csr is not None:
        m_indptr = np.ascontiguousarray(mask_csr.indptr, dtype=np.int64)
        m_indices = np.ascontiguousarray(mask_csr.indices, dtype=np.int32)
        keep.extend([m_indptr, m_indices])
        a.has_mask = 1
        a.mask_indptr, a.mask_indices = m_indptr.ctypes.data, m_indices.ctypes.data
    a.sym_upper = int(bool(sym_upper))
    a.max_dist = -1 if max_dist is None else int(max_dist)
    a.full = int(bool(full))
    a.pval = int(bool(pval))
    a.trim_to_max_dist = int(bool(trim_to_max_dist))
    # canonical CSR (sorted rows): the library measures the diagonal extent itself
    a.sig_dmin, a.sig_dmax = -(2 ** 31), -1
    if on_device:
        a.sig_dmin, a.sig_dmax = csr.diag_range
    a.kernel = _kernel_desc(kernel, tsvd, keep)
    a.missing_tol = float(missing_tol)
    a.device = _device_index() if device is None else int(device)
    a.raw_xcorr = int(bool(raw_xcorr))
    a.xcorr_threshold = float(threshold)
    if out_rows is not None:
        a.out_row0, a.out_row1 = int(out_rows[0]), int(out_rows[1])
    return a, keep


class _PinnedOwner:
    """Keeps the library's pinned result buffers alive for as long as a numpy view of
    them is (zero-copy hand-over to scipy); returns them to the pool afterwards."""

    def __init__(self, res):
        self._res = res

    def __del__(self):
        try:
            _lib.load().cs_result_free(C.byref(self._res))
        except Exception:
            pass


def _view(owner, ptr, ctype, dtype, n):
    """numpy view of n elements of a library-owned pinned buffer."""
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    buf = (ctype * n).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=n)
    # frombuffer keeps `buf` alive; tie the owner to it as well
    buf._owner = owner
    return arr


def _result_to_csr(res, shape, pval):
    """cs_csr_result -> (corr, pvals) scipy CSR matrices.  The value arrays are
    zero-copy views of the pinned buffers the device wrote into; the buffers go
    back to the library's pool when the matrices are garbage collected."""
    n = int(res.nnz)
    owner = _PinnedOwner(res)
    indptr_o = _view(owner, res.indptr, C.c_int64, np.int64, shape[0] + 1)
    idx = _view(owner, res.indices, C.c_int32, np.int32, n)
    val = _view(owner, res.data, C.c_double, np.float64, n)
    # scipy wants one index dtype for both structure arrays
    if n < 2 ** 31:
        indptr_o = indptr_o.astype(np.int32)
    else:
        idx = idx.astype(np.int64)
    corr = sp.csr_matrix(shape, dtype=np.float64)
    corr.indptr, corr.indices, corr.data = indptr_o, idx, val
    pvals = None
    if pval:
        pv = _view(owner, res.log10p, C.c_double, np.float64, n)
        # own structure arrays: callers compact the two matrices independently (det:267)
        p_indptr = _view(owner, res.p_indptr, C.c_int64, np.int64, shape[0] + 1)
        p_idx = _view(owner, res.p_indices, C.c_int32, np.int32, n)
        if n < 2 ** 31:
            p_indptr = p_indptr.astype(np.int32)
        else:
            p_idx = p_idx.astype(np.int64)
        pvals = sp.csr_matrix(shape, dtype=np.float64)
        pvals.indptr, pvals.indices, pvals.data = p_indptr, p_idx, pv
    return corr, pvals


def _run_host(csr, kernel, mask_csr, max_dist, sym_upper, full, missing_tol, tsvd, pval,
              raw_xcorr=False, threshold=1e-4, trim_to_max_dist=False, geometry=None):
    """One call of cs_normxcorr2_host -> (corr csr, log10 p csr or None)."""
    lib = _lib.load()
    a, keep = _build_args(csr, kernel, mask_csr, max_dist, sym_upper, full, missing_tol, tsvd,
                          pval, raw_xcorr, threshold, trim_to_max_dist, geometry=geometry)
    res = _lib.CsrResult()
    _lib.check(lib.cs_normxcorr2_host(C.byref(a), C.byref(res)))
    last_call_stats.clear()
    last_call_stats.update(ms_h2d=res.ms_h2d, ms_kernels=res.ms_kernels, ms_d2h=res.ms_d2h,
                           n_windows=int(res.n_windows), nnz=int(res.nnz),
                           h2d_bytes=int(res.h2d_bytes), d2h_bytes=int(res.d2h_bytes))
    del keep
    return _result_to_csr(res, csr.shape, pval)


def _check_kernel_shape(kernel):
    km, kn = kernel.shape
    if km % 2 == 0 or kn % 2 == 0:
        # even kernels break the reference too (inconsistent shapes after zero_pad_sparse,
        # det:720-722); fail early with a clear message
        raise ValueError("kernel dimensions must be odd")


def _validate(signal, kernel, missing_mask):
    """The argument checks of det:871-889, in the reference's order."""
    if missing_mask is not None:
        if not sp.issparse(missing_mask):
            raise ValueError("Missing mask must be a sparse matrix.")
        if not signal.shape == missing_mask.shape:
            raise ValueError("Signal and missing mask do not have the same shape")
        if missing_mask.dtype != bool:
            raise ValueError(f"Missing mask dtype is {missing_mask.dtype}. Should be bool.")
        if min(kernel.shape) >= max(signal.shape):
            raise ValueError("cannot have kernel bigger than signal")
    if sp.issparse(kernel):
        raise ValueError("cannot handle kernel in sparse format")
    kernel = np.asarray(kernel, dtype=np.float64)
    if not (kernel.std() > 0):
        raise ValueError("Cannot have flat kernel.")
    _check_kernel_shape(kernel)
    return kernel


def _mask_forms(missing_mask, sym_upper):
    """(pixel mask CSR, geometry): the geometric form when the mask is one
    make_missing_mask can build (two bit vectors reach the device instead of the pixel
    list), else the canonical CSR pattern."""
    if missing_mask is None:
        return None, None
    geometry = None if MASK_FORM == "pixels" else preproc.mask_geometry(missing_mask, sym_upper)
    if geometry is not None:
        return None, geometry
    return _mask_csr(missing_mask), None


def _mask_csr(missing_mask):
    if missing_mask is None:
        return None
    mask_csr = missing_mask.tocsr()
    if mask_csr.nnz and not mask_csr.data.all():
        mask_csr = mask_csr.copy()
        mask_csr.eliminate_zeros()
    if not mask_csr.has_canonical_format:
        mask_csr = mask_csr.copy()
        mask_csr.sum_duplicates()
    return mask_csr


def xcorr2(signal, kernel, threshold=1e-4, tsvd=None):
    """Cross-correlation of a 2-D signal with a dense kernel (det:595-624).

    Same shape as the signal, zero margins of half a kernel, values below
    `threshold` in magnitude set to zero.  Sparse in -> csr_matrix out, dense in
    -> ndarray out."""
    if isinstance(kernel, tuple):
        left, right = kernel
        if left.shape[1] != right.shape[0]:
            raise ValueError("Kernel factorisation is invalid")
        kernel = np.asarray(left) @ np.asarray(right)
        tsvd = None
    if sp.issparse(kernel):
        raise ValueError("cannot handle kernel in sparse format")
    kernel = np.asarray(kernel, dtype=np.float64)
    _check_kernel_shape(kernel)
    dense_in = not sp.issparse(signal)
    csr = _canonical_csr(np.asarray(signal) if dense_in else signal, np.float64)
    out, _ = _run_host(csr, kernel, None, None, False, False, 0.75, tsvd, False,
                       raw_xcorr=True, threshold=threshold)
    return out.toarray() if dense_in else out


def normxcorr2(
    signal,
    kernel,
    max_dist=None,
    sym_upper=False,
    full=False,
    missing_mask=None,
    missing_tol=0.75,
    tsvd=None,
    pval=False,
    *,
    trim_to_max_dist=False,
):
    """Pearson correlation of every window of `signal` with `kernel`
    (det:807-914).  Arguments, return values and ValueErrors follow the
    reference; `trim_to_max_dist` is an extension that skips the scores beyond
    `max_dist` which pattern_detector throws away anyway (det:270).

    Returns (corr, log10_pvals): csr_matrix for a sparse signal, ndarray for a
    dense one; log10_pvals is None unless pval=True."""
    kernel = _validate(signal, kernel, missing_mask)
    dense_in = not sp.issparse(signal)
    csr = _canonical_csr(np.asarray(signal) if dense_in else signal, np.float64)
    mask_csr, geometry = _mask_forms(missing_mask, sym_upper)
    corr, pvals = _run_host(csr, kernel, mask_csr, max_dist, sym_upper, full, missing_tol, tsvd,
                            pval, trim_to_max_dist=trim_to_max_dist, geometry=geometry)
    if dense_in:
        return corr.toarray(), (pvals.toarray() if pvals is not None else None)
    return corr, pvals


# --------------------------------------------------------------------------- callers of the hot path
def validate_patterns(coords, matrix, conv_mat, detectable_bins, kernel_matrix, drop=True,
                      zero_tol=0.3, missing_tol=0.75):
    """Windows around pattern coordinates, their validation and score (det:18-155).

    One CUDA warp gathers each window from the CSR matrix (csrc/gather.cu) instead of
    the reference's Python loop.  Returns (DataFrame[bin1, bin2, score], windows)."""
    import pandas as pd
    from .. import _cuda
    t = _cuda.require_cuda()
    lib = _lib.load()
    coords = np.asarray(coords).reshape(-1, 2)
    P = coords.shape[0]
    csr = sp.csr_matrix(matrix, dtype=np.float64)
    if not csr.has_sorted_indices:
        csr = csr.copy()
        csr.sort_indices()
    km, kn = kernel_matrix.shape
    windows = np.full((P, km, kn), np.nan)
    valid = np.zeros(P, dtype=bool)
    if P:
        vr = np.zeros(csr.shape[0], dtype=np.uint8)
        vr[np.asarray(detectable_bins[0], dtype=np.int64)] = 1
        vc = np.zeros(csr.shape[1], dtype=np.uint8)
        vc[np.asarray(detectable_bins[1], dtype=np.int64)] = 1
        d = dict(indptr=_cuda.to_device(csr.indptr, np.int64), indices=_cuda.to_device(csr.indices, np.int32),
                 data=_cuda.to_device(csr.data, np.float64), vr=_cuda.to_device(vr), vc=_cuda.to_device(vc),
                 coords=_cuda.to_device(coords, np.int32))
        d_win = _cuda.empty(P * km * kn, t.float64)
        d_ok = _cuda.empty(P, t.uint8)
        g = _lib.GatherArgs()
        g.rows, g.cols = csr.shape
        g.d_indptr, g.d_indices, g.d_data = d["indptr"].data_ptr(), d["indices"].data_ptr(), d["data"].data_ptr()
        g.d_valid_row, g.d_valid_col = d["vr"].data_ptr(), d["vc"].data_ptr()
        g.win_h, g.win_w = km, kn
        g.zero_tol, g.missing_tol = float(zero_tol), float(missing_tol)
        _lib.check(lib.cs_window_gather(C.byref(g), _cuda.ptr(d["coords"]), P, _cuda.ptr(d_win),
                                        _cuda.ptr(d_ok), _cuda.stream_ptr()))
        windows = d_win.cpu().numpy().reshape(P, km, kn)
        valid = d_ok.cpu().numpy().astype(bool)
    score = np.full(P, np.nan)
    if valid.any():
        cm = sp.csr_matrix(conv_mat)
        score[valid] = np.asarray(cm[coords[valid, 0], coords[valid, 1]]).ravel()
    table = pd.DataFrame({"bin1": coords[:, 0], "bin2": coords[:, 1], "score": score})
    if drop:
        return table.loc[valid, :], windows[valid]
    return table, windows


_sessions = {}
# running totals over pattern_detector calls (bench.py --config 4 / 5 report the device share)
detector_totals = {"calls": 0, "device_ms": 0.0, "wall_ms": 0.0}


def _detector_session():
    """One device-resident session per (process, device): pattern_detector is called once per
    sub-matrix and kernel (cli:738-792); allocating ~1.5 GB of device buffers per call would
    cost more than the call."""
    import threading
    from ..session import Session
    dev = _device_index()
    s = _sessions.get(dev)
    if s is None:
        s = _sessions[dev] = (Session(dev), threading.Lock())
    return s


def pattern_detector(contact_map, kernel_config, kernel_matrix, coords=None, dump=None, full=False,
                     tsvd=None):
    """Detect (or, with `coords`, quantify) one pattern kernel on one sub-matrix
    (det:177-345).  `contact_map` is anything with the attributes of
    contacts_map.ContactMap that the reference reads: matrix, detectable_bins, max_dist,
    inter (and name when dumping).

    The sub-matrix stays in HBM for the whole call: normxcorr2, the thresholding of
    pick_foci, the window gather / validation and the score and p-value lookups run on
    the device; only candidate pixels, windows and the final table come back.
    Returns (DataFrame[bin1, bin2, score, pvalue], windows) or (None, None)."""
    import pandas as pd
    from ..session import Session, records_to_numpy
    kernel_matrix = np.asarray(kernel_matrix, dtype=np.float64)
    km, kn = kernel_matrix.shape
    kh, kw = (km - 1) // 2, (kn - 1) // 2
    quantify = coords is not None
    # a ContactMap preprocessed on the device (contacts_map.py) hands its CSR over in HBM
    signal = getattr(contact_map, "device_csr", None)
    if signal is None:
        signal = contact_map.matrix
    shape = signal.shape
    if min(shape) <= max(kernel_matrix.shape):           # det:237-238
        return None, None
    inter = bool(contact_map.inter)
    geometry = None
    if full:
        # det:241-250 builds the pixel mask of the missing bins; the device only needs its
        # ingredients (two bit vectors and the flagged diagonals)
        geometry = preproc.missing_geometry(
            shape, contact_map.detectable_bins[0], contact_map.detectable_bins[1],
            max_dist=contact_map.max_dist, sym_upper=not inter)
    import time as _time
    sess, lock = _detector_session()
    lock.acquire()
    _t0 = _time.perf_counter()
    try:
        # the detector reads the score image only (foci, lookups at coordinates): upload and
        # kernels overlap slab by slab, the CSR compaction and the p-values of every stored
        # score are skipped unless dumped
        sess.upload(signal, kernel_matrix, max_dist=contact_map.max_dist,
                    sym_upper=not inter, full=full, mask_geometry=geometry, tsvd=tsvd, pval=True,
                    missing_tol=kernel_config["max_perc_undetected"] / 100, run_scores=not dump)
        if dump:
            sess.run(compact=True)
        dmax = 2 ** 30 if inter else int(contact_map.max_dist)
        dmin = -(2 ** 30) if inter else 0
        if dump:
            import pathlib
            conv, _ = sess.download()
            sp.save_npz(pathlib.Path(dump) / f"{contact_map.name}_03_normxcorr2", conv)
            if not inter:
                sp.save_npz(pathlib.Path(dump) / f"{contact_map.name}_04_diag_trim",
                            preproc.diag_trim(conv.tocsr(), contact_map.max_dist))
        if not quantify:
            # det:277-283: pixels above the threshold -> 4-connected foci -> one local maximum
            # per focus, on the device (csrc/scores.cu); only the foci leave it
            thr = float(kernel_config["pearson"])
            if dump:
                # the dump wants the labelled matrix: candidates to the host, host labelling
                import pathlib
                cap = 1 << 20
                while True:
                    rec, n = sess.candidates(thr, dmin, dmax, cap=cap)
                    if n < cap:
                        break
                    cap *= 4
                cand = records_to_numpy(rec, n)
                cmat = sp.coo_matrix((cand["score"].astype(np.float64), (cand["row"], cand["col"])),
                                     shape=shape)
                coords, foci_mat = pick_foci(cmat, thr)
                if coords is None:
                    return None, None
                sp.save_npz(pathlib.Path(dump) / f"{contact_map.name}_05_foci", foci_mat.tocsr())
            else:
                foci = sess.foci(thr, dmin, dmax, min_size=2)
                if len(foci) == 0:
                    return None, None
                coords = np.stack([foci["row"], foci["col"]], axis=1)
        coords = np.array(coords, dtype=np.int64).reshape(-1, 2)
        if not inter and kernel_config["max_dist"] == 0:
            # det:311-315: 1-D patterns sit on the diagonal of the padded map
            coords[:, 0] = coords[:, 1] + ((kw - kh) if full else 0)
        windows, valid, score, logp = sess.validate(
            coords, contact_map.detectable_bins[0], contact_map.detectable_bins[1], inter,
            kernel_config["max_perc_zero"] / 100, kernel_config["max_perc_undetected"] / 100, dmax)
    finally:
        detector_totals["calls"] += 1
        detector_totals["device_ms"] += float(sess.stats.get("ms_total", 0.0))
        detector_totals["wall_ms"] += 1e3 * (_time.perf_counter() - _t0)
        lock.release()  # the session (and its device buffers) is reused by the next call
    score = np.where(valid, score, np.nan)
    table = pd.DataFrame({"bin1": coords[:, 0], "bin2": coords[:, 1], "score": score,
                          "pvalue": 10.0 ** logp})
    if not quantify:                                     # det:328: drop in detect mode only
        table = table.loc[valid, :].reset_index(drop=True)
        windows = windows[valid]
    return table, windows


# --------------------------------------------------------------------------- foci (host)
def label_foci(matrix):
    """4-connected component labelling of the non-zero pixels (det:459-554).
    Labels start at 1 and are numbered by the first pixel of each focus in
    row-major order."""
    coo = sp.coo_matrix(sp.csr_matrix(matrix))
    n = coo.nnz
    order = np.lexsort((coo.col, coo.row))
    row, col = coo.row[order].astype(np.int64), coo.col[order].astype(np.int64)
    width = int(coo.shape[1]) + 1
    key = row * width + col
    idx = np.arange(n)
    # right neighbours are adjacent in row-major order; lower neighbours by key lookup
    right = np.flatnonzero((row[1:] == row[:-1]) & (col[1:] == col[:-1] + 1))
    below_pos = np.searchsorted(key, key + width)
    below_ok = (below_pos < n)
    below_ok[below_ok] = key[below_pos[below_ok]] == (key + width)[below_ok]
    a = np.concatenate([right, idx[below_ok]])
    b = np.concatenate([right + 1, below_pos[below_ok]])
    graph = sp.coo_matrix((np.ones(len(a), dtype=np.int8), (a, b)), shape=(n, n))
    num, comp = sp.csgraph.connected_components(graph, directed=False)
    # number components by first appearance in row-major order
    first = np.full(num, n, dtype=np.int64)
    np.minimum.at(first, comp, idx)
    rank = np.empty(num, dtype=np.int64)
    rank[np.argsort(first)] = np.arange(num)
    labels = rank[comp] + 1
    out = sp.coo_matrix((labels, (row, col)), shape=coo.shape)
    return num, out


def filter_foci(foci_mat, min_size=2):
    """Drop foci of fewer than min_size pixels (det:557-592)."""
    data = foci_mat.data
    ids, sizes = np.unique(data, return_counts=True)
    small = ids[sizes < min_size]
    data = np.where(np.isin(data, small), 0, data)
    out = foci_mat.copy()
    out.data = data
    out.eliminate_zeros()
    return int((sizes >= min_size).sum()), out


def pick_foci(mat_conv, pearson, min_size=2):
    """Local maxima of the thresholded correlation map, one per focus
    (det:387-456).  Returns (coords, labelled matrix) or (None, None)."""
    cand = sp.coo_matrix(mat_conv).copy()
    keep = (cand.data >= pearson) & (cand.data != 0)
    cand = sp.coo_matrix((np.ones(keep.sum()), (cand.row[keep], cand.col[keep])), shape=cand.shape)
    if cand.nnz == 0:
        return None, None
    _, labelled = label_foci(cand)
    num, labelled = filter_foci(labelled, min_size=min_size)
    if num == 0:
        return None, None
    scores = np.asarray(sp.csr_matrix(mat_conv)[labelled.row, labelled.col]).ravel()
    coords = np.zeros((num, 2), dtype=int)
    order = np.argsort(labelled.data, kind="stable")
    lab_sorted = labelled.data[order]
    starts = np.flatnonzero(np.r_[True, lab_sorted[1:] != lab_sorted[:-1]])
    ends = np.r_[starts[1:], len(order)]
    for k, (s, e) in enumerate(zip(starts, ends)):
        members = order[s:e]                       # row-major order within the focus
        best = members[np.argmax(scores[members])]  # first maximum, as np.argmax
        coords[k] = labelled.row[best], labelled.col[best]
    return coords, labelled


def pileup_patterns(pattern_windows):
    """Arithmetic mean of a stack of windows, ignoring NaN (det:158-174)."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        return np.nanmean(np.asarray(pattern_windows, dtype=np.float64), axis=0)


def remove_neighbours(patterns, win_size=8):
    """Patterns closer than win_size pixels (on both axes) to a better-scoring one are
    dropped (det:348-384): boolean mask over the rows of `patterns` (DataFrame with bin1,
    bin2, score), greedy by decreasing score.  Same result as the reference's O(P^2) Python
    loop; candidates are looked up through a grid of win_size x win_size buckets."""
    n = len(patterns)
    keep = np.ones(n, dtype=bool)
    if n == 0:
        return keep
    b1 = np.asarray(patterns.bin1, dtype=np.int64)
    b2 = np.asarray(patterns.bin2, dtype=np.int64)
    score = np.asarray(patterns.score, dtype=np.float64)
    # descending score, NaN last, ties in input order (what sort_values gives)
    order = np.argsort(-np.nan_to_num(score, nan=-np.inf), kind="stable")
    w = max(int(win_size), 1)
    cell = {}
    for i in range(n):
        cell.setdefault((b1[i] // w, b2[i] // w), []).append(i)
    dead = np.zeros(n, dtype=bool)
    for i in order:
        if dead[i]:
            continue
        c1, c2 = b1[i] // w, b2[i] // w
        for a in (c1 - 1, c1, c1 + 1):
            for b in (c2 - 1, c2, c2 + 1):
                for j in cell.get((a, b), ()):
                    if j != i and abs(b1[j] - b1[i]) < w and abs(b2[j] - b2[i]) < w:
                        dead[j] = True
    keep[dead] = False
    return keep
