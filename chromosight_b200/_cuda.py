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


class DeviceCSR:
    """A canonical CSR matrix whose column indices (int32) and values (float64) live in HBM as
    torch tensors; the row pointers (int64) stay on the host, which plans the call from them.
    What the device-side preprocessing of a sub-matrix (contacts_map.ContactMap.create_mat)
    hands to Session.upload without a host round trip.  `diag_range` = (min, max) of
    col - row over the stored entries."""

    def __init__(self, shape, indptr, d_indices, d_data, diag_range):
        self.shape = (int(shape[0]), int(shape[1]))
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        self.d_indices = d_indices
        self.d_data = d_data
        self.diag_range = (int(diag_range[0]), int(diag_range[1]))
        self.nnz = int(self.indptr[-1])
        self.dtype = np.dtype(np.float64)

    def to_scipy(self, drop_zeros=True):
        """Host copy as a scipy CSR matrix (explicit zeros removed)."""
        import scipy.sparse as sp
        m = sp.csr_matrix((self.d_data.cpu().numpy(), self.d_indices.cpu().numpy(), self.indptr.copy()),
                          shape=self.shape)
        m.has_canonical_format = True
        if drop_zeros:
            m.eliminate_zeros()
        return m
