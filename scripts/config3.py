#!/usr/bin/env python3
"""BASELINE.json config 3 end to end: detect borders on the synthetic 200k map with the three border
kernels resized to 9x9 (--win-size 9), min-dist 0, max-dist 2 Mb, through pattern_detector (upload,
normxcorr2, foci, window validation on the device)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from chromosight_b200 import kernels
from chromosight_b200.utils import detection as cud, preprocessing as cup

n = int(os.environ.get("N", 200000)); D = 200
cfg = dict(kernels.borders)
ks = [cup.resize_kernel(np.asarray(k, dtype=np.float64), factor=9 / 17) for k in cfg["kernels"]]
cfg["kernels"] = ks
cfg["max_dist"] = 2_000_000   # the CLI override of config 3 (--max-dist 2000000): scan 200 bins
k = ks[0].shape[0]


class Map:
    inter = False
    name = "synthetic"


raw, detect = bench.raw_map(n, D, k, 0)
mat = cup.detrend(raw, detectable_bins=detect, max_dist=D + k, max_val=10)
mat = cup.diag_trim(mat.tocsr(), D + k)
mat.data[np.isnan(mat.data)] = 0
mat.eliminate_zeros()
m = Map(); m.matrix = mat; m.max_dist = D; m.detectable_bins = (detect, detect)
cud.pattern_detector(m, cfg, ks[0], full=True)   # warm-up
out = []
for i, kern in enumerate(ks):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    table, windows = cud.pattern_detector(m, cfg, kern, full=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out.append({"kernel": i, "shape": list(kern.shape), "seconds": dt,
                "patterns": 0 if table is None else len(table)})
print(json.dumps({"config": 3, "n_bins": n, "max_dist_bins": D, "pearson": cfg["pearson"], "runs": out}))
