"""ctypes binding of libchromosight_b200.so (the C ABI of include/chromosight_b200.h).

There is no CPU fallback: if the library is missing or was built for another
architecture, importing the hot-path functions raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CHROMOSIGHT_B200_LIB selects another build of the same ABI (the -DCS_ABLATE build that
# scripts/k1_ablate.sh times)
LIB_PATH = os.environ.get("CHROMOSIGHT_B200_LIB") or os.path.join(_HERE, "libchromosight_b200.so")

ABI_VERSION = 3  # CS_ABI_VERSION of include/chromosight_b200.h
CS_OK = 0
CS_ERR_INVALID = -1
CS_ERR_CUDA = -2
CS_ERR_NOMEM = -3
CS_ERR_MASKED_SIGNAL = -4


class Layout(C.Structure):
    _fields_ = [
        ("rows", C.c_int32), ("cols", C.c_int32),
        ("dlo", C.c_int32), ("dhi", C.c_int32),
        ("pitch", C.c_int32), ("dense", C.c_int32),
        ("n_elems", C.c_int64),
    ]


class KernelDesc(C.Structure):
    _fields_ = [
        ("kh", C.c_int32), ("kw", C.c_int32),
        ("k_corr", C.c_void_p), ("k_mask", C.c_void_p), ("k2_mask", C.c_void_p),
        ("k_sum", C.c_double), ("k2_sum", C.c_double),
        ("k_mean", C.c_double), ("k_std", C.c_double),
    ]


class GeoMask(C.Structure):
    _fields_ = [
        ("d_row_bits", C.c_void_p), ("d_col_bits", C.c_void_p),
        ("mask_dlo", C.c_int32), ("mask_dhi", C.c_int32),
        ("mat_y0", C.c_int32), ("mat_y1", C.c_int32),
        ("mat_x0", C.c_int32), ("mat_x1", C.c_int32),
        ("margin_mode", C.c_int32), ("top_x1", C.c_int32), ("right_y0", C.c_int32),
        ("strip_dlo", C.c_int32), ("strip_dhi", C.c_int32),
        ("fill_value", C.c_float),
    ]


class PearsonOpts(C.Structure):
    _fields_ = [
        ("mask_mode", C.c_int32),
        ("missing_tol", C.c_double),
        ("xcorr_threshold", C.c_double),
        ("raw_xcorr", C.c_int32),
        ("nobs_full", C.c_int32),
        ("tile_rows", C.c_int32),
        ("out_row_shift", C.c_int32),
        ("out_col_shift", C.c_int32),
        ("nmiss_bytes", C.c_int32),
        ("geo", GeoMask),
    ]


class Candidate(C.Structure):
    _fields_ = [("row", C.c_int32), ("col", C.c_int32), ("score", C.c_float), ("log10p", C.c_float)]


CANDIDATE_DTYPE = np.dtype([("row", "<i4"), ("col", "<i4"), ("score", "<f4"), ("log10p", "<f4")])
# cs_focus records (24 B)
FOCUS_DTYPE = np.dtype([("first_row", "<i4"), ("first_col", "<i4"), ("row", "<i4"), ("col", "<i4"),
                        ("score", "<f4"), ("size", "<i4")])


class CsrResult(C.Structure):
    _fields_ = [
        ("nnz", C.c_int64),
        ("rows", C.c_int32), ("cols", C.c_int32),
        ("indptr", C.c_void_p), ("indices", C.c_void_p),
        ("data", C.c_void_p), ("log10p", C.c_void_p),
        ("p_indptr", C.c_void_p), ("p_indices", C.c_void_p),
        ("ms_h2d", C.c_double), ("ms_kernels", C.c_double), ("ms_d2h", C.c_double),
        ("n_windows", C.c_int64),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
    ]


class Normxcorr2Args(C.Structure):
    _fields_ = [
        ("rows", C.c_int32), ("cols", C.c_int32),
        ("indptr", C.c_void_p), ("indices", C.c_void_p), ("data", C.c_void_p),
        ("has_mask", C.c_int32),
        ("mask_indptr", C.c_void_p), ("mask_indices", C.c_void_p),
        ("miss_row", C.c_void_p), ("miss_col", C.c_void_p),
        ("mask_dlo", C.c_int32), ("mask_dhi", C.c_int32),
        ("sym_upper", C.c_int32), ("max_dist", C.c_int32),
        ("full", C.c_int32), ("pval", C.c_int32),
        ("trim_to_max_dist", C.c_int32),
        ("sig_dmin", C.c_int32), ("sig_dmax", C.c_int32),
        ("kernel", KernelDesc),
        ("missing_tol", C.c_double),
        ("device", C.c_int32),
        ("raw_xcorr", C.c_int32),
        ("xcorr_threshold", C.c_double),
        ("device_payload", C.c_int32),
        ("out_row0", C.c_int32), ("out_row1", C.c_int32),
    ]


class GatherArgs(C.Structure):
    _fields_ = [
        ("rows", C.c_int32), ("cols", C.c_int32),
        ("d_indptr", C.c_void_p), ("d_indices", C.c_void_p), ("d_data", C.c_void_p),
        ("d_valid_row", C.c_void_p), ("d_valid_col", C.c_void_p),
        ("win_h", C.c_int32), ("win_w", C.c_int32),
        ("pad_rows", C.c_int32), ("pad_cols", C.c_int32),
        ("det_shift_row", C.c_int32), ("det_shift_col", C.c_int32),
        ("nan_subdiag", C.c_int32),
        ("zero_tol", C.c_double), ("missing_tol", C.c_double),
    ]


class RunStats(C.Structure):
    _fields_ = [
        ("ms_fill", C.c_double), ("ms_pearson", C.c_double),
        ("ms_compact", C.c_double), ("ms_total", C.c_double),
        ("n_windows", C.c_int64), ("nnz", C.c_int64), ("launches", C.c_int64),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
    ]


# name -> (restype, argtypes); also the list of symbols the header declares
_P = C.c_void_p
_PROTOS = {
    "cs_version": (C.c_int, []),
    "cs_last_error": (C.c_char_p, []),
    "cs_launch_count": (C.c_int64, []),
    "cs_layout_band": (C.c_int, [C.POINTER(Layout), C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "cs_layout_band_padded": (C.c_int, [C.POINTER(Layout), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int32]),
    "cs_layout_dense": (C.c_int, [C.POINTER(Layout), C.c_int32, C.c_int32]),
    "cs_image_fill_f32": (C.c_int, [C.POINTER(Layout), _P, _P, _P, _P, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_int32, C.c_int32, _P, _P, C.POINTER(GeoMask),
                                     C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "cs_pearson_f32": (C.c_int, [C.POINTER(Layout), _P, C.POINTER(KernelDesc), C.POINTER(PearsonOpts),
                                  C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.POINTER(Layout), _P, _P, _P]),
    "cs_pearson_tile_rows": (C.c_int, [C.POINTER(Layout), C.POINTER(KernelDesc), C.POINTER(PearsonOpts),
                                        C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.POINTER(C.c_int32)]),
    "cs_pearson_plan": (C.c_int, [C.POINTER(Layout), C.POINTER(KernelDesc), C.POINTER(PearsonOpts),
                                   C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "cs_pixels_lex_sorted": (C.c_int, [_P, _P, C.c_int64]),
    "cs_pixels_inter_index": (C.c_int64, [_P, _P, C.c_int64, _P, C.c_int32, _P, _P]),
    "cs_band_csr_from_pixels": (C.c_int64, [_P, _P, _P, C.c_int32, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int64,
                                            _P, _P, _P, C.c_int32]),
    "cs_scan_scratch": (C.c_int64, [C.c_int32]),
    "cs_scores_count": (C.c_int, [C.POINTER(Layout), _P, C.c_int32, C.c_int32, _P,
                                   C.POINTER(C.c_int64), _P]),
    "cs_scores_emit": (C.c_int, [C.POINTER(Layout), _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  _P, _P, _P, _P, _P]),
    "cs_scores_candidates": (C.c_int, [C.POINTER(Layout), _P, _P, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_float, _P, C.c_int64, _P, C.POINTER(C.c_int64),
                                        _P]),
    "cs_window_gather": (C.c_int, [C.POINTER(GatherArgs), _P, C.c_int64, _P, _P, _P]),
    "cs_scores_lookup": (C.c_int, [C.POINTER(Layout), _P, _P, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_int32, _P, C.c_int64, _P, _P, _P]),
    "cs_session_validate": (C.c_int, [C.c_void_p, _P, C.c_int64, _P, _P, C.c_int32, C.c_double,
                                       C.c_double, C.c_int32, _P, _P, _P, _P]),
    "cs_distance_law": (C.c_int, [_P, _P, _P, C.c_int32, _P, C.c_int32, _P, _P, _P, _P]),
    "cs_detrend_apply": (C.c_int, [_P, _P, _P, _P, C.c_int32, _P, C.c_int32, C.c_double, _P]),
    "cs_expand_rows": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, C.c_int32]),
    "cs_normxcorr2_host": (C.c_int, [C.POINTER(Normxcorr2Args), C.POINTER(CsrResult)]),
    "cs_result_free": (None, [C.POINTER(CsrResult)]),
    "cs_session_create": (C.c_int, [C.c_int32, C.POINTER(C.c_void_p)]),
    "cs_session_destroy": (None, [C.c_void_p]),
    "cs_session_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cs_session_upload": (C.c_int, [C.c_void_p, C.POINTER(Normxcorr2Args)]),
    "cs_session_run": (C.c_int, [C.c_void_p, C.POINTER(RunStats)]),
    "cs_session_run_enqueue": (C.c_int, [_P]),
    "cs_session_wait": (C.c_int, [_P, C.POINTER(RunStats)]),
    "cs_session_run_scores": (C.c_int, [C.c_void_p, C.POINTER(RunStats)]),
    "cs_session_upload_run_scores": (C.c_int, [C.c_void_p, C.POINTER(Normxcorr2Args), C.POINTER(RunStats)]),
    "cs_foci_work_bytes": (C.c_int64, [C.POINTER(Layout)]),
    "cs_scores_foci": (C.c_int, [C.POINTER(Layout), _P, C.c_int32, C.c_int32, C.c_double, C.c_int32, _P, _P,
                                  C.c_int64, _P, C.POINTER(C.c_int64), _P]),
    "cs_session_foci": (C.c_int, [C.c_void_p, C.c_double, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int64,
                                   C.POINTER(C.c_int64)]),
    "cs_session_candidates": (C.c_int, [C.c_void_p, C.c_float, C.c_int32, C.c_int32, _P, C.c_int64,
                                         _P, C.POINTER(C.c_int64)]),
    "cs_session_download": (C.c_int, [C.c_void_p, C.POINTER(CsrResult)]),
}
EXPORTED_SYMBOLS = tuple(_PROTOS)

_lib = None


class BackendError(RuntimeError):
    """The CUDA library is missing or failed; there is no CPU fallback."""


def load():
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BackendError(
            f"{LIB_PATH} not found: build it with `python chromosight_b200/csrc/build.py` "
            "(there is no CPU fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.cs_version() != ABI_VERSION:
        raise BackendError("libchromosight_b200.so ABI version mismatch")
    _lib = lib
    return lib


def last_error():
    return load().cs_last_error().decode("utf-8", "replace")


def check(rc):
    """Translate a cs_status into the exception the reference would raise."""
    if rc == CS_OK:
        return
    msg = last_error()
    if rc in (CS_ERR_INVALID, CS_ERR_MASKED_SIGNAL):
        raise ValueError(msg)
    if rc == CS_ERR_NOMEM:
        raise MemoryError(msg)
    raise BackendError(msg)


def launch_count():
    return int(load().cs_launch_count())
