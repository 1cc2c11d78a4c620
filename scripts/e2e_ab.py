#!/usr/bin/env python3
"""A/B of the end-to-end host path on the GPU box: one process per setting (the library reads
its knobs once), the bench workload, wall time per detection.normxcorr2 call.
usage: python scripts/e2e_ab.py            (driver: runs the settings below)
       python scripts/e2e_ab.py child TAG  (one measurement)"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SETTINGS = [
    ("narrow default", {}),
    ("narrow 2 thr", {"CS_EXPAND_THREADS": "2"}),
    ("narrow 4 thr", {"CS_EXPAND_THREADS": "4"}),
    ("narrow 12 thr", {"CS_EXPAND_THREADS": "12"}),
    ("wide (r1 path)", {"CS_WIDE_RESULT": "1"}),
    ("narrow trace", {"CS_TRACE": "1"}),
    ("trace plainmemcpy", {"CS_TRACE": "1", "CS_STAGE_PLAIN_MEMCPY": "1"}),
    ("trace wide", {"CS_TRACE": "1", "CS_WIDE_RESULT": "1"}),
]

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    import bench
    from chromosight_b200 import kernels, _cuda
    from chromosight_b200.utils import detection as cud, preprocessing as cup
    n, D = 200000, 200
    kernel = np.asarray(kernels.loops["kernels"][0], dtype=np.float64); k = kernel.shape[0]
    raw, detect = bench.raw_map(n, D, k, 0)
    mat = cup.detrend(raw, detectable_bins=detect, max_dist=D + k, max_val=10)
    mat, mask = bench.finish_map(mat, detect, D, k, cup.diag_trim, cup.make_missing_mask)
    kw = bench.call_kwargs(D)
    trace = os.environ.pop("CS_TRACE", None)
    for label, m in (("pageable", mat), ("pinned", _cuda.pin_sparse(mat))):
        for _ in range(3):
            r, p = cud.normxcorr2(m, kernel, missing_mask=mask, **kw)
        torch.cuda.synchronize()
        ts = []
        for _ in range(6):
            t0 = time.perf_counter(); r, p = cud.normxcorr2(m, kernel, missing_mask=mask, **kw); ts.append(time.perf_counter() - t0)
        print(f"  {sys.argv[2]:16s} {label:9s} ms/call: {[round(1e3 * t, 1) for t in ts]}  span {cud.last_call_stats['ms_kernels']:.1f} ms", flush=True)
        if trace and label == "pageable":
            os.environ["CS_TRACE"] = "1"
            r, p = cud.normxcorr2(m, kernel, missing_mask=mask, **kw)
            os.environ.pop("CS_TRACE")
else:
    only = os.environ.get("AB_ONLY")
    for tag, env in SETTINGS:
        if only and only not in tag:
            continue
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child", tag], env=e, timeout=600)
