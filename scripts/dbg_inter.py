#!/usr/bin/env python3
"""Debug: one inter-chromosomal sub-matrix of the config-5 genome through pattern_detector,
with the session's stage times.  usage: python scripts/dbg_inter.py [rows] [cols]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from chromosight_b200 import kernels, synthetic
from chromosight_b200.contacts_map import HicGenome
from chromosight_b200.utils import detection as cud
ms, ns = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (24000, 20000)
clr = synthetic.genome_cool([ms, ns], binsize=10_000, n_diags=217, seed=10, inter_density=1e-4, density_floor=1.0)
cfg = dict(kernels.loops); cfg["kernels"] = [np.array(k) for k in cfg["kernels"]]
hg = HicGenome(clr, inter=True, kernel_config=cfg); hg.normalize(); hg.make_sub_matrices()
for _, row in hg.sub_mats.iterrows():
    cm = row.contact_map
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        cm.create_mat()
        torch.cuda.synchronize(); t1 = time.perf_counter()
        table, wins = cud.pattern_detector(cm, cfg, cfg["kernels"][0], full=True)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        cm.destroy_mat()
    sess, _ = cud._detector_session()
    print(cm.name, "inter" if cm.inter else "intra", cm.shape, "create_mat %.1f ms, pattern_detector %.1f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1)),
          "patterns", 0 if table is None else len(table), {k: round(v, 2) if isinstance(v, float) else v for k, v in sess.stats.items()})
