"""Device-resident form of the hot-path call (cs_session_* of the C ABI).

`chromosight.utils.detection.normxcorr2` is one synchronous host-to-host call;
iterated detection (cli:730-792) and quantify re-run the same sub-matrix with a
new kernel, and a benchmark wants the inputs resident in HBM.  A Session splits
the call into upload / run / download (+ candidates: the thresholding of
pick_foci, det:417-421, on the device)."""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _cuda, _lib
from .utils import detection as _det


class Session:
    def __init__(self, device=None, use_torch_stream=True):
        t = _cuda.require_cuda()
        self._lib = _lib.load()
        self.device = int(t.cuda.current_device() if device is None else device)
        self._h = C.c_void_p()
        _lib.check(self._lib.cs_session_create(self.device, C.byref(self._h)))
        self._use_torch_stream = use_torch_stream
        self.shape = None
        self.pval = False
        self.stats = {}

    def close(self):
        if self._h:
            self._lib.cs_session_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _bind_stream(self):
        if self._use_torch_stream:
            t = _cuda.torch()
            with t.cuda.device(self.device):
                st = t.cuda.current_stream().cuda_stream
            # stream 0 is the legacy default stream: a valid choice, but NULL means
            # "library stream" in the ABI, so only bind real streams
            _lib.check(self._lib.cs_session_set_stream(self._h, C.c_void_p(st) if st else None))

    def upload(self, signal, kernel, max_dist=None, sym_upper=False, full=False, missing_mask=None,
               missing_tol=0.75, tsvd=None, pval=False, trim_to_max_dist=False, mask_geometry=None,
               run_scores=False, out_rows=None):
        """Plan a normxcorr2 call (same arguments, det:807-817) and copy its inputs to HBM.
        `mask_geometry` = preprocessing.missing_geometry(...) stands for the mask
        make_missing_mask would build, without building it.  run_scores=True also computes the
        scores (as run(compact=False) would), overlapping the upload of large inputs with the
        kernels slab by slab.  out_rows=(a, b) scores the rows a <= row < b only (the owned rows
        of a row slab, rowslab.py); the rest of the result stays empty."""
        kernel = np.asarray(kernel, dtype=np.float64)
        if isinstance(signal, _cuda.DeviceCSR):
            # entries already in HBM; order the library stream behind the stream that made them
            csr = signal
            if not (self._use_torch_stream and _cuda.torch().cuda.current_stream().cuda_stream):
                _cuda.torch().cuda.current_stream().synchronize()
            _det._validate(signal, kernel, missing_mask)
        else:
            _det._validate(signal, kernel, missing_mask)
            csr = _det._canonical_csr(signal, np.float64)
        mask_csr = None
        if mask_geometry is None:
            mask_csr, mask_geometry = _det._mask_forms(missing_mask, sym_upper)
        a, keep = _det._build_args(csr, kernel, mask_csr, max_dist, sym_upper, full, missing_tol,
                                   tsvd, pval, trim_to_max_dist=trim_to_max_dist, device=self.device,
                                   geometry=mask_geometry, out_rows=out_rows)
        self._bind_stream()
        if run_scores:
            st = _lib.RunStats()
            _lib.check(self._lib.cs_session_upload_run_scores(self._h, C.byref(a), C.byref(st)))
            self.stats = {f: getattr(st, f) for f, _ in _lib.RunStats._fields_}
        else:
            _lib.check(self._lib.cs_session_upload(self._h, C.byref(a)))
        del keep
        self.shape = csr.shape
        self.kernel_shape = kernel.shape
        self.pval = bool(pval)
        return self

    def run(self, compact=True, wait=True):
        """fill -> Pearson -> CSR compaction (+ p-values of every stored score) on the device;
        returns the run statistics.  compact=False stops after the Pearson kernel: candidates,
        foci and validate only read the score image (download compacts on demand).
        wait=False only enqueues the run (re-runs of one upload): the next call that
        synchronises -- candidates(), wait() -- checks it; returns None."""
        self._bind_stream()
        if not wait and compact:
            _lib.check(self._lib.cs_session_run_enqueue(self._h))
            return None
        st = _lib.RunStats()
        fn = self._lib.cs_session_run if compact else self._lib.cs_session_run_scores
        _lib.check(fn(self._h, C.byref(st)))
        self.stats = {f: getattr(st, f) for f, _ in _lib.RunStats._fields_}
        return self.stats

    def wait(self):
        """Wait for a run enqueued with run(wait=False); the statistics of the last run."""
        self._bind_stream()
        st = _lib.RunStats()
        _lib.check(self._lib.cs_session_wait(self._h, C.byref(st)))
        self.stats = {f: getattr(st, f) for f, _ in _lib.RunStats._fields_}
        return self.stats

    def download(self):
        """(corr, log10 p-values or None) of the last run as scipy CSR matrices."""
        self._bind_stream()
        res = _lib.CsrResult()
        _lib.check(self._lib.cs_session_download(self._h, C.byref(res)))
        self.stats["ms_d2h"] = res.ms_d2h
        return _det._result_to_csr(res, self.shape, self.pval)

    def candidates(self, threshold, dmin=0, dmax=2 ** 30, cap=1 << 22, out=None):
        """Pixels with score >= threshold on diagonals dmin..dmax of the last run as a
        DEVICE tensor of (row, col, score, log10p) records (16 B each, _lib.CANDIDATE_DTYPE)
        -- the send buffer of the final all-gather.  Returns (tensor, count)."""
        t = _cuda.torch()
        self._bind_stream()
        dev = t.device("cuda", self.device)
        if out is None:
            out = t.empty((cap, 4), dtype=t.int32, device=dev)
        cap = out.shape[0]
        # (the library zeroes the counter on its own stream: no torch-side fill that could
        # land after it when the session runs on the library stream)
        cnt = t.empty(1, dtype=t.int64, device=dev)
        n = C.c_int64(0)
        _lib.check(self._lib.cs_session_candidates(self._h, C.c_float(threshold), int(dmin),
                                                   int(min(dmax, 2 ** 30)), _cuda.ptr(out), cap,
                                                   _cuda.ptr(cnt), C.byref(n)))
        return out, int(min(n.value, cap))


    def foci(self, threshold, dmin=0, dmax=2 ** 30, min_size=2, cap=1 << 16):
        """pick_foci (det:387-456) on the device: 4-connected foci of the pixels with score >=
        threshold on diagonals dmin..dmax, at least min_size pixels each.  Returns a structured
        array (_lib.FOCUS_DTYPE) in the reference's focus order (by first pixel, row-major); (row,
        col) of a record is the focus' local maximum."""
        self._bind_stream()
        while True:
            rec = np.zeros(cap, dtype=_lib.FOCUS_DTYPE)
            n = C.c_int64(0)
            _lib.check(self._lib.cs_session_foci(self._h, C.c_double(threshold), int(max(dmin, -(2 ** 30))),
                                                 int(min(dmax, 2 ** 30)), int(min_size),
                                                 rec.ctypes.data, cap, C.byref(n)))
            if n.value <= cap:
                return rec[: n.value]
            cap = int(n.value)

    def validate(self, coords, valid_rows, valid_cols, inter, zero_tol, missing_tol, score_dmax):
        """validate_patterns (det:18-155) on the uploaded matrix and the last scores.

        coords : int array [P, 2] of unpadded (row, col); valid_rows / valid_cols : indices of
        the detectable bins.  Returns (windows [P, kh, kw] float64 with NaN windows for invalid
        patterns, valid bool [P], score float64 [P], log10 p float64 [P])."""
        coords = np.ascontiguousarray(coords, dtype=np.int32).reshape(-1, 2)
        P = coords.shape[0]
        km, kn = self.kernel_shape
        windows = np.empty((P, km, kn), dtype=np.float64)
        valid = np.zeros(P, dtype=np.uint8)
        score = np.zeros(P, dtype=np.float64)
        logp = np.zeros(P, dtype=np.float64)
        vr = np.zeros(self.shape[0], dtype=np.uint8)
        vr[np.asarray(valid_rows, dtype=np.int64)] = 1
        vc = np.zeros(self.shape[1], dtype=np.uint8)
        vc[np.asarray(valid_cols, dtype=np.int64)] = 1
        self._bind_stream()
        if P:
            _lib.check(self._lib.cs_session_validate(
                self._h, coords.ctypes.data, P, vr.ctypes.data, vc.ctypes.data, int(bool(inter)),
                float(zero_tol), float(missing_tol), int(min(score_dmax, 2 ** 30)),
                windows.ctypes.data, valid.ctypes.data, score.ctypes.data, logp.ctypes.data))
        return windows, valid.astype(bool), score, logp


def records_to_numpy(tensor, count):
    """Candidate records (device or host int32[*, 4] tensor) -> structured numpy array."""
    a = tensor[:count].cpu().numpy()
    return a.view(_lib.CANDIDATE_DTYPE).reshape(-1)
