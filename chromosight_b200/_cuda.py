"""Device-memory plumbing on top of PyTorch (tensors as raw device buffers, the
current CUDA stream as the launch stream).  PyTorch is only the allocator and
stream provider here; all arithmetic happens in libchromosight_b200.so."""
import ctypes as C

import numpy as np

from . import _lib

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def require_cuda():
    t = torch()
    if not t.cuda.is_available():
        raise _lib.BackendError(
            "chromosight_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    _lib.load()
    return t


def stream_ptr():
    return C.c_void_p(torch().cuda.current_stream().cuda_stream)


def to_device(arr, dtype=None, device=None):
    """numpy array -> contiguous CUDA tensor (pageable source, synchronous copy)."""
    t = torch()
    a = np.ascontiguousarray(arr, dtype=dtype)
    return t.from_numpy(a).to(device or "cuda")


def empty(n, dtype, device=None):
    return torch().empty(int(n), dtype=dtype, device=device or "cuda")


def ptr(tensor):
    return C.c_void_p(tensor.data_ptr()) if tensor is not None else C.c_void_p(0)


def pin_sparse(mat):
    """CSR copy of a scipy sparse matrix whose index and value arrays live in page-locked
    host memory (torch pinned tensors).  The library recognises pinned arrays
    (cudaPointerGetAttributes) and DMAs them straight to HBM instead of going through its
    staging buffers -- the input side of the end-to-end contract."""
    import scipy.sparse as sp
    t = torch()
    csr = sp.csr_matrix(mat)
    if not csr.has_canonical_format:
        csr = csr.copy()
        csr.sum_duplicates()

    def pin(a):
        return t.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

    out = sp.csr_matrix((pin(csr.data), pin(csr.indices), pin(csr.indptr)), shape=csr.shape, copy=False)
    out.has_canonical_format = True
    if hasattr(mat, "_cs_geometry"):  # a mask of make_missing_mask stays recognisable
        out._cs_geometry = mat._cs_geometry
    return out
