#!/usr/bin/env python3
"""Fixtures for the global host filters mirrored in chromosight_b200 (remove_neighbours
det:348-384, pileup_patterns det:158-174, fdr_correction stats:7-40), produced by the
UNMODIFIED reference:

    python tests/golden/make_golden_hostfilters.py
"""
import os
import sys
import warnings

import numpy as np
import pandas as pd

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

import chromosight.utils.detection as cud  # noqa: E402
import chromosight.utils.stats as cus  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")


def main():
    rng = np.random.default_rng(12)
    f = {}
    for i, (n, span, win) in enumerate([(1, 50, 4), (60, 120, 5), (500, 300, 8), (300, 40, 3)]):
        t = pd.DataFrame({"bin1": rng.integers(0, span, n), "bin2": rng.integers(0, span, n),
                          "score": np.round(rng.random(n), 3)})   # rounded: ties occur
        f[f"rn{i}_bin1"], f[f"rn{i}_bin2"], f[f"rn{i}_score"] = t.bin1.values, t.bin2.values, t.score.values
        f[f"rn{i}_win"] = np.int64(win)
        f[f"rn{i}_mask"] = np.asarray(cud.remove_neighbours(t, win_size=win))
    w = rng.random((9, 5, 7))
    w[2, 1, 1] = np.nan
    w[:, 0, 0] = np.nan
    f["pile_in"], f["pile_out"] = w, cud.pileup_patterns(w)
    p = np.concatenate([rng.random(40) ** 3, [0.0, 1.0, 0.05, 0.05]])
    f["fdr_in"], f["fdr_out"] = p, cus.fdr_correction(p)
    np.savez_compressed(os.path.join(OUT, "hostfilters.npz"), **f)
    print("wrote hostfilters.npz", {k: v.shape for k, v in f.items() if k.endswith("mask")})


if __name__ == "__main__":
    main()
