#!/usr/bin/env python3
"""Golden fixture for the real-data plumbing case (BASELINE.json configs[0]: detect loops on
data_test/example.cool), produced in the build container by the UNMODIFIED reference:

    python tests/golden/make_golden_cool.py

`cooler` is not installed here, so the reference's contacts_map module is imported with
chromosight_b200.cool.CoolFile standing in for `cooler.Cooler` (the file is parsed by our
minimal HDF5 reader; its tables are cross-checked against the attributes cooler wrote into
the file: nbins, nnz, sum).  Everything after the file access is the reference's own code:
ContactMap.create_mat (balance, detrend, trim) and pattern_detector, per chromosome, loops
preset.  Stored: the file's tables (so that tests do not need the .cool), the preprocessed
sub-matrices, the pattern tables and windows.
"""
import json
import os
import sys
import types
import warnings

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

from chromosight_b200.cool import CoolFile  # noqa: E402

fake = types.ModuleType("cooler")
fake.Cooler = CoolFile
sys.modules["cooler"] = fake

import chromosight.kernels as ck  # noqa: E402
import chromosight.utils.contacts_map as rcm  # noqa: E402
import chromosight.utils.detection as cud  # noqa: E402
import chromosight.utils.preprocessing as cup  # noqa: E402

from make_golden import coo_fields  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
COOL = "/root/reference/data_test/example.cool"


def main():
    clr = CoolFile(COOL)
    assert clr.info["nbins"] == clr.shape[0] and clr.info["nnz"] == len(clr.pixels()[:])
    assert clr.info["sum"] == clr.pixels()[:]["count"].sum()
    cfg = dict(ck.loops)
    kernel = np.array(cfg["kernels"][0])
    bins = clr.bins()[:]
    pix = clr.pixels()[:]
    f = {
        "chrom_names": np.array(clr.chromnames), "chrom_sizes": np.asarray(clr.chromsizes.values),
        "bin_chrom": np.asarray(bins.chrom.cat.codes, dtype=np.int32),
        "bin_start": bins.start.values, "bin_end": bins.end.values, "bin_weight": bins.weight.values,
        "pix_bin1": pix.bin1_id.values, "pix_bin2": pix.bin2_id.values, "pix_count": pix["count"].values,
        "binsize": np.int64(clr.binsize), "kernel": kernel,
        "config": np.array(json.dumps({k: v for k, v in cfg.items() if k != "kernels"})),
    }
    max_dist = max(cfg["max_dist"] // clr.binsize, 1)
    largest = kernel.shape[0]
    d = np.flatnonzero(np.isfinite(bins.weight.values))
    total = 0
    for chrom in clr.chromnames:
        s, e = clr.extent(chrom)
        det = (d[(d >= s) & (d < e)] - s, d[(d >= s) & (d < e)] - s)
        cm = rcm.ContactMap(clr, extent=[(s, e), (s, e)], name=f"{chrom}-{chrom}", detectable_bins=det,
                            inter=False, max_dist=max_dist, largest_kernel=largest, use_norm=True)
        cm.create_mat()
        f.update(coo_fields(f"{chrom}_matrix", cm.matrix))
        res, windows = cud.pattern_detector(cm, cfg, kernel, full=True)
        mask = cup.make_missing_mask(cm.matrix.shape, det[0], det[1], max_dist=max_dist, sym_upper=True)
        conv, _ = cud.normxcorr2(cm.matrix.tocsr(), kernel, max_dist=max_dist, sym_upper=True, full=True,
                                 missing_mask=mask, pval=True,
                                 missing_tol=cfg["max_perc_undetected"] / 100)
        conv.data[np.isnan(conv.data)] = 0
        conv = cup.diag_trim(conv.tocsr(), max_dist).tocsr()
        n = 0 if res is None else len(res)
        total += n
        f[f"{chrom}_n"] = np.int64(n)
        if n:
            b1, b2 = np.asarray(res.bin1, dtype=int), np.asarray(res.bin2, dtype=int)
            f[f"{chrom}_bin1"], f[f"{chrom}_bin2"] = b1, b2
            f[f"{chrom}_pvalue"] = np.asarray(res.pvalue, dtype=float)
            f[f"{chrom}_score"] = np.asarray(conv[b1, b2]).ravel()
            f[f"{chrom}_windows"] = windows
        print(f"  {chrom}: {cm.matrix.shape[0]} bins, {cm.matrix.nnz} pixels, {n} loops")
    f["max_dist"] = np.int64(max_dist)
    np.savez_compressed(os.path.join(OUT, "cool_example_loops.npz"), **f)
    print("total", total, "->", os.path.join(OUT, "cool_example_loops.npz"),
          os.path.getsize(os.path.join(OUT, "cool_example_loops.npz")) // 1024, "KB")


if __name__ == "__main__":
    main()
