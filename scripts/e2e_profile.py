#!/usr/bin/env python3
"""Where the end-to-end time of detection.normxcorr2 goes on the GPU box: cProfile of the
host-facing call on the bench workload + raw pinned PCIe copy rates for reference."""
import cProfile, pstats, io, os, sys, time
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
for _ in range(3):
    r, p = cud.normxcorr2(mat, kernel, missing_mask=mask, **kw)
torch.cuda.synchronize()
ts = []
for _ in range(6):
    t0 = time.perf_counter(); r, p = cud.normxcorr2(mat, kernel, missing_mask=mask, **kw); ts.append(time.perf_counter() - t0)
print("wall ms per call:", [round(1e3 * t, 2) for t in ts], "stats", cud.last_call_stats)
pr = cProfile.Profile(); pr.enable()
for _ in range(4):
    r, p = cud.normxcorr2(mat, kernel, missing_mask=mask, **kw)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(25); print(s.getvalue()[:6000])
# raw PCIe rates
nb = 1 << 30
h = torch.empty(nb, dtype=torch.uint8).pin_memory(); d = torch.empty(nb, dtype=torch.uint8, device="cuda")
for name, a, b in (("H2D", d, h), ("D2H", h, d)):
    for _ in range(2): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 4
    print(f"{name} pinned 1 GiB: {nb / dt / 1e9:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(); d2 = torch.empty_like(d); h2 = torch.empty(nb, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"duplex 1 GiB each way: {dt * 1e3:.1f} ms -> {nb / dt / 1e9:.1f} GB/s per direction")
print("cpus", os.cpu_count())
# host memcpy rate
a = np.empty(nb // 8, dtype=np.float64); b = np.ones(nb // 8, dtype=np.float64)
t0 = time.perf_counter(); a[:] = b; print(f"host memcpy 1 thread: {nb / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
