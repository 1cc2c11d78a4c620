"""One large chromosome over several GPUs: contiguous row slabs.

The reference never splits a chromosome (one pool worker per sub-matrix, cli:738-755); for the
1/2/4/8-GPU metric on a single 200k-bin map the rows are cut into `world` slabs instead:

* every rank owns the output rows [r0, r1) and receives the square sub-matrix of rows and
  columns [in0, in1) with a halo of k rows above and D + 3k below (windows and the frame
  geometry of frame_missing_mask, pre:404-498, then see the same pixels as in the whole map);
* the distance law (pre:129-197) is global: every rank sums its OWNED rows per diagonal, ONE
  all-reduce of 2 x (max_dist + 1) numbers makes the law, then every rank detrends its slab
  (pre:256-310) with it;
* candidate pixels (not foci) of the owned rows are gathered once, so that foci straddling a
  slab boundary are labelled once, globally (pick_foci, det:387-456, on rank 0).
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _cuda, _lib, sharding


def slab_plan(n_rows, world_size, kernel_size, max_dist):
    """[(r0, r1, in0, in1)] per rank: owned output rows and the rows / columns of the square
    input sub-matrix."""
    bounds = np.linspace(0, n_rows, world_size + 1).round().astype(int)
    k, below = int(kernel_size), int(max_dist) + 3 * int(kernel_size)
    return [(int(bounds[g]), int(bounds[g + 1]), max(int(bounds[g]) - k, 0),
             min(int(bounds[g + 1]) + below, n_rows)) for g in range(world_size)]


def owned_rows(raw, r0, r1):
    """CSR of the whole shape whose rows outside [r0, r1) are empty: the rows a rank adds to the
    distance-law sums."""
    raw = raw.tocsr()
    n = raw.shape[0]
    a, b = raw.indptr[r0], raw.indptr[r1]
    indptr = np.concatenate([np.zeros(r0, np.int64), raw.indptr[r0:r1 + 1].astype(np.int64) - a,
                             np.full(n - r1, b - a, np.int64)])
    return sp.csr_matrix((raw.data[a:b], raw.indices[a:b], indptr), shape=raw.shape)


def law_sums(raw_owned, detectable_bins, max_dist):
    """Per-diagonal (sum, count) of the strictly positive pixels with both bins detectable
    (pre:178-188), on the device (K0a).  Returns two torch tensors of length n_diags."""
    t = _cuda.require_cuda()
    lib = _lib.load()
    csr = raw_owned.tocsr()
    n = csr.shape[0]
    n_diags = int(min(n, max_dist + 1))
    flags = np.zeros(n, dtype=np.uint8)
    flags[np.asarray(detectable_bins)] = 1
    d = (_cuda.to_device(csr.indptr, np.int64), _cuda.to_device(csr.indices, np.int32),
         _cuda.to_device(csr.data, np.float64), _cuda.to_device(flags))
    d_sum, d_cnt = _cuda.empty(n_diags, t.float64), _cuda.empty(n_diags, t.int64)
    d_law = _cuda.empty(n, t.float64)
    _lib.check(lib.cs_distance_law(_cuda.ptr(d[0]), _cuda.ptr(d[1]), _cuda.ptr(d[2]), n, _cuda.ptr(d[3]),
                                   n_diags, _cuda.ptr(d_sum), _cuda.ptr(d_cnt), _cuda.ptr(d_law),
                                   _cuda.stream_ptr()))
    return d_sum, d_cnt


def law_from_sums(total_sum, total_cnt, n):
    """The law detrend divides by (pre:289: diagonals without pixels -> 0), length n."""
    law = np.zeros(n)
    m = len(total_sum)
    ok = total_cnt > 0
    law[:m][ok] = total_sum[ok] / total_cnt[ok]
    return law


def detrend_with_law(sub, law, max_val=10):
    """pre:300-309 with a given law: data / law[|row - col|], values >= max_val -> 1 (device)."""
    t = _cuda.require_cuda()
    lib = _lib.load()
    csr = sp.csr_matrix(sub, dtype=np.float64)
    if csr.nnz == 0:
        return csr
    d = (_cuda.to_device(csr.indptr, np.int64), _cuda.to_device(csr.indices, np.int32),
         _cuda.to_device(csr.data, np.float64), _cuda.to_device(law, np.float64))
    out = _cuda.empty(csr.nnz, t.float64)
    _lib.check(lib.cs_detrend_apply(_cuda.ptr(d[0]), _cuda.ptr(d[1]), _cuda.ptr(d[2]), _cuda.ptr(out),
                                    csr.shape[0], _cuda.ptr(d[3]), len(law),
                                    C.c_double(-1.0 if max_val is None else float(max_val)),
                                    _cuda.stream_ptr()))
    return sp.csr_matrix((out.cpu().numpy(), csr.indices.copy(), csr.indptr.copy()), shape=csr.shape)


def global_law(raw, detectable_bins, max_dist, r0, r1, group=None):
    """The whole map's distance law from every rank's owned rows: ONE all-reduce of the
    per-diagonal sums and counts (world size 1: no collective)."""
    import torch
    import torch.distributed as dist
    d_sum, d_cnt = law_sums(owned_rows(raw, r0, r1), detectable_bins, max_dist)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        buf = torch.cat([d_sum, d_cnt.to(torch.float64)])   # counts < 2^53: exact in float64
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        m = d_sum.numel()
        d_sum, d_cnt = buf[:m], buf[m:].round().to(torch.int64)
    return law_from_sums(d_sum.cpu().numpy(), d_cnt.cpu().numpy(), raw.shape[0])


def slab_inputs(raw, detectable_bins, law, plan, max_dist, kernel_size, max_val=10):
    """The detrended, trimmed square sub-matrix of one slab and its detectable bins, as
    ContactMap.create_mat would build them for a chromosome (cm:607-624)."""
    from .utils import preprocessing as cup
    r0, r1, in0, in1 = plan
    sub = raw.tocsr()[in0:in1, in0:in1]
    mat = detrend_with_law(sub, law[: in1 - in0], max_val)
    mat = cup.diag_trim(mat.tocsr(), max_dist + kernel_size)
    mat.data[np.isnan(mat.data)] = 0
    mat.eliminate_zeros()
    det = np.asarray(detectable_bins)
    det = det[(det >= in0) & (det < in1)] - in0
    return mat, det


def owned_candidates(records, plan):
    """Candidate records (slab coordinates) of the owned rows, shifted to chromosome
    coordinates."""
    r0, r1, in0, _ = plan
    rec = records[(records["row"] >= r0 - in0) & (records["row"] < r1 - in0)].copy()
    rec["row"] += in0
    rec["col"] += in0
    return rec


def merge_sorted(parts):
    """Candidate records of all slabs in row-major order (what one run over the whole map
    yields after sorting)."""
    rec = np.concatenate(parts) if parts else np.zeros(0, dtype=_lib.CANDIDATE_DTYPE)
    order = np.lexsort((rec["col"], rec["row"]))
    return rec[order]


def foci_of_candidates(records, shape, pearson):
    """pick_foci (det:387-456) on gathered candidate pixels: the global labelling step of the
    row-slab path (rank 0)."""
    from .utils import detection as cud
    if len(records) == 0:
        return None
    cmat = sp.coo_matrix((records["score"].astype(np.float64), (records["row"], records["col"])), shape=shape)
    coords, _ = cud.pick_foci(cmat, pearson)
    return coords


__all__ = ["slab_plan", "owned_rows", "law_sums", "law_from_sums", "detrend_with_law", "global_law",
           "slab_inputs", "owned_candidates", "merge_sorted", "foci_of_candidates", "sharding"]
