#!/usr/bin/env python3
"""Golden fixture for the whole `chromosight detect` chain on data_test/example.cool
(cli:625-860), produced in the build container with the UNMODIFIED reference's functions:

    python tests/golden/make_golden_cli.py

The docopt CLI, cooler and matplotlib are not installed here, so cmd_detect is driven by hand:
per chromosome the reference's ContactMap.create_mat + pattern_detector (with
chromosight_b200.cool.CoolFile standing in for cooler.Cooler, as in make_golden_cool.py), then
the reference's own remove_neighbours / pileup_patterns / fdr_correction in the order of
cli:757-848.  pandas 3 turns the reference's chained assignment `validated_coords.score[i] = ...`
(det:134) into a no-op, which would leave every score NaN; the score column is therefore read
from the reference's trimmed correlation map at the coordinates (what det:134 assigns).

Four command lines (docs/notebooks/plot_output.ipynb:13-15 and `chromosight test`, cli:185-199):
    loops_default   detect example.cool                      (TEST_LOG: "89 patterns detected")
    loops_nb        detect example.cool -m8000 -M50000 -p0.35   (docs .../example_loops.tsv, 59 rows)
    borders         detect example.cool --pattern borders       (docs .../example_borders.tsv, 57 rows)
    hairpins        detect example.cool --pattern hairpins      (docs .../example_hairpins.tsv, 55 rows)
Stored per case: the final table (bins, kernel_id, iteration, score, pvalue, qvalue) and the
bins of the TSV the reference's repository holds for it (written by an older chromosight
version; kept for the comparison the test reports).
"""
import os
import sys
import types
import warnings

import numpy as np
import pandas as pd

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
from chromosight.utils.stats import fdr_correction  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
COOL = "/root/reference/data_test/example.cool"
DOCS = "/root/reference/docs/notebooks/detect"

CASES = {
    "loops_default": ("loops", {}, None),
    "loops_nb": ("loops", {"min_dist": 8000, "max_dist": 50000, "pearson": 0.35}, "example_loops.tsv"),
    "borders": ("borders", {}, "example_borders.tsv"),
    "hairpins": ("hairpins", {}, "example_hairpins.tsv"),
}


def detect(clr, cfg):
    binsize = clr.binsize
    bins = clr.bins()[:]
    d = np.flatnonzero(np.isfinite(bins.weight.values))
    largest = max(np.array(k).shape[0] for k in cfg["kernels"])
    max_dist = max(cfg["max_dist"] // binsize, 1)                       # cm:168-172
    all_coords, all_windows = [], []
    for kernel_id, kernel in enumerate(cfg["kernels"]):
        kernel = np.array(kernel, dtype=float)
        for it in range(cfg["max_iterations"]):
            tabs, wins = [], []
            for chrom in clr.chromnames:
                s, e = clr.extent(chrom)
                det = (d[(d >= s) & (d < e)] - s, d[(d >= s) & (d < e)] - s)
                cm = rcm.ContactMap(clr, extent=[(s, e), (s, e)], name=f"{chrom}-{chrom}", detectable_bins=det,
                                    inter=False, max_dist=max_dist, largest_kernel=largest, use_norm=True)
                cm.create_mat()
                res, windows = cud.pattern_detector(cm, cfg, kernel, full=True)
                if res is None:
                    continue
                # det:134 under pandas >= 3 (see the module docstring)
                mask = cup.make_missing_mask(cm.matrix.shape, det[0], det[1], max_dist=max_dist, sym_upper=True)
                conv, _ = cud.normxcorr2(cm.matrix.tocsr(), kernel, max_dist=max_dist, sym_upper=True,
                                         full=True, missing_mask=mask, pval=True,
                                         missing_tol=cfg["max_perc_undetected"] / 100)
                conv.data[np.isnan(conv.data)] = 0
                conv = cup.diag_trim(conv.tocsr(), max_dist).tocsr()
                res = res.reset_index(drop=True)
                b1, b2 = np.asarray(res.bin1, dtype=int), np.asarray(res.bin2, dtype=int)
                res["score"] = np.asarray(conv[b1, b2]).ravel()
                res["bin1"] = b1 + s
                res["bin2"] = b2 + s
                tabs.append(res)
                wins.append(windows)
            if not wins:
                break
            wins = np.concatenate(wins, axis=0)
            tab = pd.concat(tabs, axis=0).reset_index(drop=True)
            tab["kernel_id"] = kernel_id
            tab["iteration"] = it
            all_coords.append(tab)
            all_windows.append(wins)
            kernel = cud.pileup_patterns(wins)                           # cli:791
    if not all_coords:
        return None
    tab = pd.concat(all_coords, axis=0).reset_index(drop=True)
    sep = max(int(cfg["min_separation"] // binsize), 1)                  # cli:808-811
    keep = cud.remove_neighbours(tab, win_size=sep)                      # cli:814-816
    tab = tab.loc[keep, :].reset_index(drop=True)
    start1, start2 = bins.start.values[tab.bin1.values], bins.start.values[tab.bin2.values]
    chrom1, chrom2 = bins.chrom.values[tab.bin1.values], bins.chrom.values[tab.bin2.values]
    near = (np.asarray(chrom1) == np.asarray(chrom2)) & (np.abs(start2 - start1) < cfg["min_dist"])  # cli:836-839
    tab = tab.loc[~near, :]
    tab = tab.loc[~tab.pvalue.isnull(), :].reset_index(drop=True)        # cli:844-846
    tab["qvalue"] = fdr_correction(tab["pvalue"])                        # cli:848
    return tab


def main():
    clr = CoolFile(COOL)
    f = {}
    for name, (pattern, override, tsv) in CASES.items():
        cfg = dict(getattr(ck, pattern))
        cfg.update(override)
        tab = detect(clr, cfg)
        for c in ("bin1", "bin2", "kernel_id", "iteration"):
            f[f"{name}_{c}"] = tab[c].values.astype(np.int64)
        for c in ("score", "pvalue", "qvalue"):
            f[f"{name}_{c}"] = tab[c].values.astype(np.float64)
        f[f"{name}_override"] = np.array(repr(override))
        f[f"{name}_pattern"] = np.array(pattern)
        msg = f"  {name}: {len(tab)} patterns"
        if tsv:
            held = pd.read_csv(os.path.join(DOCS, tsv), sep="\t")
            f[f"{name}_held_bin1"], f[f"{name}_held_bin2"] = held.bin1.values, held.bin2.values
            f[f"{name}_held_score"] = held.score.values
            ours = set(zip(tab.bin1, tab.bin2))
            theirs = set(zip(held.bin1, held.bin2))
            msg += f"; repository TSV {tsv}: {len(held)} rows, {len(ours & theirs)} shared"
        print(msg)
    path = os.path.join(OUT, "cli_example.npz")
    np.savez_compressed(path, **f)
    print("->", path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
