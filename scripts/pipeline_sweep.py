#!/usr/bin/env python3
"""Slab-count sweep and per-slab timeline (CS_TRACE) of the host-facing normxcorr2 call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from chromosight_b200 import kernels
from chromosight_b200.utils import detection as cud, preprocessing as cup

n = int(os.environ.get("N", 200000)); D = 200
kernel = np.asarray(kernels.loops["kernels"][0], dtype=np.float64); k = kernel.shape[0]
raw, detect = bench.raw_map(n, D, k, 0)
mat = cup.detrend(raw, detectable_bins=detect, max_dist=D + k, max_val=10)
mat, mask = bench.finish_map(mat, detect, D, k, cup.diag_trim, cup.make_missing_mask)
kw = bench.call_kwargs(D)
if os.environ.get('PINNED'):
    from chromosight_b200 import _cuda
    mat, mask = _cuda.pin_sparse(mat), _cuda.pin_sparse(mask)
for slabs in [int(x) for x in os.environ.get("SLABS", "8,4,12,16,24,32,48").split(",")]:
    os.environ["CS_PIPELINE_SLABS"] = str(slabs)
    for _ in range(3):
        r, p = cud.normxcorr2(mat, kernel, missing_mask=mask, **kw)
    ts = []
    for _ in range(6):
        t0 = time.perf_counter(); r, p = cud.normxcorr2(mat, kernel, missing_mask=mask, **kw); ts.append(time.perf_counter() - t0)
    print(f"slabs {slabs:3d}: wall ms min {1e3*min(ts):.2f} med {1e3*float(np.median(ts)):.2f}  span {cud.last_call_stats['ms_kernels']:.2f}", flush=True)
os.environ["CS_PIPELINE_SLABS"] = os.environ.get("TRACE_SLABS", "8")
os.environ["CS_TRACE"] = "1"
sys.stderr.flush()
r, p = cud.normxcorr2(mat, kernel, missing_mask=mask, **kw)
