#!/usr/bin/env python3
"""Where the milliseconds outside the overlapped span of detection.normxcorr2 go."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np, torch
import bench
from chromosight_b200 import kernels, _cuda, _lib
from chromosight_b200.utils import detection as cud, preprocessing as cup
n, D = 200000, 200
kernel = np.asarray(kernels.loops["kernels"][0], dtype=np.float64); k = 17
raw, detect = bench.raw_map(n, D, k, 0)
mat = cup.detrend(raw, detectable_bins=detect, max_dist=D + k, max_val=10)
mat, mask = bench.finish_map(mat, detect, D, k, cup.diag_trim, cup.make_missing_mask)
mat, mask = _cuda.pin_sparse(mat), _cuda.pin_sparse(mask)
kw = bench.call_kwargs(D)
for _ in range(3): r, p = cud.normxcorr2(mat, kernel, missing_mask=mask, **kw)
lib = _lib.load()
T = {k_: [] for k_ in ("validate", "canon", "build", "call", "result", "total")}
for _ in range(8):
    t0 = time.perf_counter()
    kk = cud._validate(mat, kernel, mask); t1 = time.perf_counter()
    csr = cud._canonical_csr(mat, np.float64); mk = cud._mask_csr(mask); t2 = time.perf_counter()
    a, keep = cud._build_args(csr, kk, mk, D, True, True, 0.5, None, True); t3 = time.perf_counter()
    res = _lib.CsrResult(); _lib.check(lib.cs_normxcorr2_host(C.byref(a), C.byref(res))); t4 = time.perf_counter()
    r, p = cud._result_to_csr(res, csr.shape, True); t5 = time.perf_counter()
    for k_, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0)): T[k_].append(v * 1e3)
print({k_: round(float(np.median(v)), 3) for k_, v in T.items()}, "span", round(res.ms_kernels, 2))
os.environ["CS_TRACE"] = "1"
r, p = cud.normxcorr2(mat, kernel, missing_mask=mask, **kw)
