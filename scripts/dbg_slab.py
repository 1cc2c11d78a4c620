#!/usr/bin/env python3
"""Debug: Pearson kernel time on one row slab of the 200k map (bench.py --scaling strong),
with and without the owned-row restriction.  usage: python scripts/dbg_slab.py [world] [rank]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from chromosight_b200 import kernels, rowslab
from chromosight_b200.session import Session
from chromosight_b200.utils import preprocessing as cup
world = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 1
kernel = np.asarray(kernels.loops["kernels"][0], dtype=np.float64); k = kernel.shape[0]
n, D = 200_000, 200
raw, detect_all = bench.raw_map(n, D, k, seed=0)
plan = rowslab.slab_plan(n, world, k, D)[rank]
law = rowslab.global_law(raw, detect_all, D + k, plan[0], plan[1])
# (single process: the law of the owned rows only -- fine for timing)
mat, detect = rowslab.slab_inputs(raw, detect_all, law, plan, D, k)
mask = cup.make_missing_mask(mat.shape, detect, detect, max_dist=D, sym_upper=True)
kw = bench.call_kwargs(D)
print("plan", plan, "slab shape", mat.shape, "nnz", mat.nnz, "geometry tag", hasattr(mask, "_cs_geometry"))
for rows in (None, (plan[0] - plan[2], plan[1] - plan[2])):
    s = Session()
    s.upload(mat, kernel, missing_mask=mask, out_rows=rows, **kw)
    for _ in range(3):
        st = s.run()
    ms = [s.run()["ms_pearson"] for _ in range(5)]
    print("out_rows", rows, "windows", st["n_windows"], "pearson ms", [round(m, 3) for m in ms], "fill", round(st["ms_fill"], 3),
          "compact", round(st["ms_compact"], 3))
    s.close()
