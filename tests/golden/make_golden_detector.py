#!/usr/bin/env python3
"""Golden fixtures for pattern_detector / validate_patterns (det:18-155, det:177-345),
produced by the UNMODIFIED reference in the build container:

    python tests/golden/make_golden_detector.py

The reference's `score` column is all-NaN under pandas 3 (chained assignment at det:134),
so the expected scores are stored as what that line means: the trimmed correlation map read
at the pattern coordinates.
"""
import json
import os
import sys
import warnings

import numpy as np
import scipy.sparse as sp

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

import chromosight.kernels as ck  # noqa: E402
import chromosight.utils.detection as cud  # noqa: E402
import chromosight.utils.preprocessing as cup  # noqa: E402

from chromosight_b200 import synthetic  # noqa: E402
from make_golden import coo_fields  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")


class DummyMap:
    """The 4-attribute stand-in for ContactMap used by the reference's tests
    (tests/test_detection.py:88-100)."""

    def __init__(self, matrix, max_dist=None, detectable_bins=None, inter=False):
        self.matrix = matrix
        self.max_dist = max_dist
        self.inter = inter
        self.detectable_bins = detectable_bins
        self.name = "dummy"


def intra_map(n, D, k, seed, missing_frac=0.03):
    raw, detect = synthetic.band_counts(n, D + k, seed=seed, missing_frac=missing_frac, max_dist=D)
    mat = cup.detrend(raw, detectable_bins=detect, max_dist=D + k, max_val=10)
    mat = cup.diag_trim(mat.tocsr(), D + k)
    mat.data[np.isnan(mat.data)] = 0
    mat.eliminate_zeros()
    return mat.tocsr(), detect


def dense_intra_map(n, D, k, seed, missing_frac=0.03, n_loops=25):
    """Fully populated band (few zero pixels, so that patterns survive max_perc_zero)
    with planted 3x3 blobs."""
    rng = np.random.default_rng(seed)
    W = D + k + 1
    d = np.arange(W)
    band = rng.poisson(60.0 / (1.0 + d) ** 0.5 + 8.0, size=(n, W)).astype(float)
    lr = rng.integers(5, n - D - 5, n_loops)
    ld = rng.integers(12, D - 8, n_loops)
    for dr in (-1, 0, 1):
        for dc in (-1, 0, 1):
            band[lr + dr, ld + dc - dr] *= 3.0
    missing = rng.random(n) < missing_frac
    rows = np.repeat(np.arange(n), W)
    cols = rows + np.tile(d, n)
    vals = band.ravel()
    keep = (cols < n) & (vals != 0)
    keep &= ~(missing[rows] | missing[np.minimum(cols, n - 1)])
    raw = sp.csr_matrix((vals[keep], (rows[keep], cols[keep])), shape=(n, n))
    detect = np.flatnonzero(~missing)
    mat = cup.detrend(raw, detectable_bins=detect, max_dist=D + k, max_val=10)
    mat = cup.diag_trim(mat.tocsr(), D + k)
    mat.data[np.isnan(mat.data)] = 0
    mat.eliminate_zeros()
    return mat.tocsr(), detect


def dense_inter_map(ms, ns, seed, missing_frac=0.04, n_blobs=12):
    rng = np.random.default_rng(seed)
    a = rng.gamma(8.0, 0.125, size=(ms, ns))
    for _ in range(n_blobs):
        r, c = rng.integers(4, ms - 4), rng.integers(4, ns - 4)
        a[r - 1:r + 2, c - 1:c + 2] *= 3.0
    mr = rng.random(ms) < missing_frac
    mc = rng.random(ns) < missing_frac
    a[mr, :] = 0
    a[:, mc] = 0
    return sp.csr_matrix(a), (np.flatnonzero(~mr), np.flatnonzero(~mc))


def run_case(name, cmap, cfg, kernel, coords=None, full=True):
    if len(sys.argv) > 1 and name not in sys.argv[1:]:   # `make_golden_detector.py NAME ...`: only these
        return
    kernel = np.array(kernel)
    res, windows = cud.pattern_detector(cmap, cfg, kernel, coords=None if coords is None else coords.copy(),
                                        full=full)
    # expected scores: conv_mat[p1, p2] of the trimmed map (what det:134 means)
    mask = None
    if full:
        mask = cup.make_missing_mask(cmap.matrix.shape, cmap.detectable_bins[0], cmap.detectable_bins[1],
                                     max_dist=cmap.max_dist, sym_upper=not cmap.inter)
    conv, _ = cud.normxcorr2(cmap.matrix.tocsr(), kernel, max_dist=cmap.max_dist, sym_upper=not cmap.inter,
                             full=full, missing_mask=mask, pval=True,
                             missing_tol=cfg["max_perc_undetected"] / 100)
    conv.data[np.isnan(conv.data)] = 0
    if not cmap.inter:
        conv = cup.diag_trim(conv.tocsr(), cmap.max_dist)
    conv = conv.tocsr()
    f = {}
    f.update(coo_fields("matrix", cmap.matrix))
    f["detect_rows"] = np.asarray(cmap.detectable_bins[0])
    f["detect_cols"] = np.asarray(cmap.detectable_bins[1])
    f["kernel"] = kernel
    f["meta"] = np.array(json.dumps({"max_dist": cmap.max_dist, "inter": bool(cmap.inter), "full": full,
                                     "config": {k: v for k, v in cfg.items() if k != "kernels"},
                                     "quantify": coords is not None, "none": res is None}))
    if coords is not None:
        f["coords_in"] = np.asarray(coords)
    if res is not None:
        b1, b2 = np.asarray(res.bin1, dtype=int), np.asarray(res.bin2, dtype=int)
        f["bin1"], f["bin2"] = b1, b2
        f["pvalue"] = np.asarray(res.pvalue, dtype=float)
        sc = np.asarray(conv[b1, b2]).ravel() if len(b1) else np.zeros(0)
        # quantify mode keeps invalid patterns with NaN windows and NaN score
        invalid = np.isnan(windows).all(axis=(1, 2)) if len(b1) else np.zeros(0, bool)
        sc = np.where(invalid, np.nan, sc)
        f["score"] = sc
        f["windows"] = windows
    np.savez_compressed(os.path.join(OUT, f"detector_{name}.npz"), **f)
    print(f"  {name}: {'None' if res is None else len(res)} patterns")


def main():
    loops = dict(ck.loops)
    borders = dict(ck.borders)
    hairpins = dict(ck.hairpins)
    # 1. detect loops, production settings
    mat, det = dense_intra_map(600, 60, 17, seed=41)
    run_case("detect_loops", DummyMap(mat, 60, (det, det)), loops, loops["kernels"][0])
    # 2. quantify given coordinates (valid, near edges, out of bounds, on missing bins)
    rng = np.random.default_rng(4)
    b1 = rng.integers(0, 600, 80)
    b2 = np.minimum(b1 + rng.integers(0, 60, 80), 599)
    coords = np.c_[b1, b2]
    coords[:4] = [[0, 0], [599, 599], [3, 40], [590, 598]]
    run_case("quantify_loops", DummyMap(mat, 60, (det, det)), loops, loops["kernels"][0], coords=coords)
    # 2b. quantify beyond max_dist and below the diagonal: the score is read from the trimmed
    # map (0 there, det:270 / det:134), the p-value from the untrimmed one (det:337-339)
    rng = np.random.default_rng(5)
    b1 = rng.integers(0, 500, 70)
    b2 = np.minimum(b1 + rng.integers(40, 95, 70), 599)
    coords = np.c_[b1, b2]
    coords[:6] = [[300, 290], [120, 100], [10, 75], [500, 585], [64, 125], [200, 262]]
    run_case("quantify_far", DummyMap(mat, 60, (det, det)), loops, loops["kernels"][0], coords=coords)
    # 3. borders: 1-D pattern (max_dist = 0 -> scan diagonals 0..1, bin1 := bin2)
    mat2, det2 = intra_map(500, 1, 17, seed=42)
    run_case("detect_borders", DummyMap(mat2, 1, (det2, det2)), borders, borders["kernels"][0])
    # 4. hairpins 15x15
    run_case("detect_hairpins", DummyMap(mat2, 1, (det2, det2)), hairpins, hairpins["kernels"][0])
    # 5. inter-chromosomal rectangle, loops_small 7x7
    imat, (vr, vc) = dense_inter_map(300, 220, seed=43)
    small = dict(ck.loops_small)
    run_case("detect_inter", DummyMap(imat.tocsr(), None, (vr, vc), inter=True), small, small["kernels"][0])
    # 6. not full (valid mode, no mask), detect
    run_case("detect_valid_mode", DummyMap(mat, 60, (det, det)), loops, loops["kernels"][0], full=False)
    # 7. nothing detected -> (None, None)
    cfg = dict(loops)
    cfg["pearson"] = 0.999
    run_case("detect_nothing", DummyMap(mat, 60, (det, det)), cfg, loops["kernels"][0])


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    main()
