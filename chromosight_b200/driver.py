"""Sharded `detect` over the sub-matrices of a genome: the loop of cmd_detect
(cli/chromosight.py:702-860) with one rank per GPU in place of the multiprocessing pool
(cli:738-755), SURVEY 8e / 8f-4.

Every rank runs `pattern_detector` (GPU) on its share of the sub-matrices; the per-sub-matrix
tables -- a few KB -- are exchanged as float64 record tensors (sharding.gather_rows: NCCL on
the device, gloo in the CPU tests; nothing is pickled) once per kernel iteration, after which
every rank holds the same global table and does the reference's global host steps (neighbour
removal, minimum distance, FDR; cli:807-848) redundantly.  Windows travel the same way, or
stay on their rank (`gather_windows=False`: only their mean, which the next iteration uses as
kernel, is all-reduced).
"""
import numpy as np
import pandas as pd

from . import sharding
from .utils import detection as cud
from .utils.stats import fdr_correction


def unit_costs(hic_genome):
    """Windows per sub-matrix: (D + 1) * n for an intra map, ms * ns for an inter map."""
    costs = []
    for _, row in hic_genome.sub_mats.iterrows():
        (s1, e1), (s2, e2) = row.contact_map.extent
        if row.contact_map.inter:
            costs.append(float(e1 - s1) * float(e2 - s2))
        else:
            D = hic_genome.max_dist if hic_genome.max_dist is not None else (e1 - s1)
            costs.append(float(min(D, e1 - s1) + 1) * float(e1 - s1))
    return costs


def _world():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), dist
    except Exception:
        pass
    return 0, 1, None


def detect_sub_matrices(hic_genome, cfg, kernel_matrix, full=True, tsvd=None, gather_windows=True):
    """One pass of `_detect_sub_mat` (cli:601-622) over all sub-matrices, sharded over the
    ranks.  Returns the list of result dicts {coords, windows, chr1, chr2} in sub-matrix
    order; the tables are identical on every rank, the windows too unless
    gather_windows=False (then `windows` is None for the sub-matrices of other ranks)."""
    rank, world, dist = _world()
    mine = sharding.partition_units(unit_costs(hic_genome), world)[rank]
    local = {}
    for u in mine:
        row = hic_genome.sub_mats.iloc[u]
        cm = row.contact_map
        cm.create_mat()
        coords, windows = cud.pattern_detector(cm, cfg, kernel_matrix, full=full, tsvd=tsvd)
        cm.destroy_mat()
        local[u] = {"coords": coords, "windows": windows, "chr1": row.chr1, "chr2": row.chr2}
    if world > 1:
        km, kn = np.asarray(kernel_matrix).shape
        have = [u for u in mine if local[u]["coords"] is not None and len(local[u]["coords"])]
        # float64 records (unit, bin1, bin2, score, pvalue): the integers are exact
        rec = np.concatenate([np.column_stack([np.full(len(local[u]["coords"]), float(u)),
                                               local[u]["coords"][["bin1", "bin2", "score", "pvalue"]]
                                               .to_numpy(dtype=np.float64)]) for u in have]) \
            if have else np.zeros((0, 5))
        parts = sharding.gather_rows(rec)
        wparts = None
        if gather_windows:
            w = np.concatenate([np.asarray(local[u]["windows"], dtype=np.float64).reshape(-1, km * kn)
                                for u in have]) if have else np.zeros((0, km * kn))
            wparts = sharding.gather_rows(w)
        out = {}
        for r, p in enumerate(parts):
            units = p[:, 0].astype(np.int64)
            for u in np.unique(units):
                sel = units == u
                srow = hic_genome.sub_mats.iloc[int(u)]
                tab = pd.DataFrame({"bin1": p[sel, 1].astype(np.int64), "bin2": p[sel, 2].astype(np.int64),
                                    "score": p[sel, 3], "pvalue": p[sel, 4]})
                wins = wparts[r][sel].reshape(-1, km, kn) if wparts is not None else \
                    (local[int(u)]["windows"] if r == rank else None)
                out[int(u)] = {"coords": tab, "windows": wins, "chr1": srow.chr1, "chr2": srow.chr2}
        for u in range(len(hic_genome.sub_mats)):
            if u not in out:
                srow = hic_genome.sub_mats.iloc[u]
                out[u] = {"coords": None, "windows": None, "chr1": srow.chr1, "chr2": srow.chr2}
        local = out
    return [local[u] for u in sorted(local)]


def _global_pileup(windows_list, shape):
    """nanmean over the windows of every rank (cli:791, det:158-174) from per-rank NaN-aware
    sums and counts: one all-reduce of 2 k^2 numbers instead of moving the windows."""
    import torch
    rank, world, dist = _world()
    tot = np.zeros(shape)
    cnt = np.zeros(shape)
    for w in windows_list:
        if w is not None and len(w):
            w = np.asarray(w, dtype=np.float64)
            tot += np.nansum(w, axis=0)
            cnt += (~np.isnan(w)).sum(axis=0)
    if world > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" \
            else torch.device("cpu")
        buf = torch.from_numpy(np.stack([tot, cnt])).to(dev)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        tot, cnt = buf.cpu().numpy()
    with np.errstate(all="ignore"):
        return tot / cnt


def detect(hic_genome, cfg, full=True, tsvd=None, gather_windows=True):
    """Pattern detection on every sub-matrix with every kernel of `cfg`, then the global
    filters (cli:720-860).  Returns (table, windows) or (None, None) when nothing is found;
    the table has the columns the reference writes (chrom1 ... qvalue).  With
    gather_windows=False the windows stay on the rank that cut them: the second value then
    holds NaN windows for the patterns of other ranks."""
    if hic_genome.sub_mats is None:
        hic_genome.make_sub_matrices()
    all_coords, all_windows = [], []
    for kernel_id, kernel_matrix in enumerate(cfg["kernels"]):
        kernel_matrix = np.asarray(kernel_matrix, dtype=np.float64)
        for it in range(cfg["max_iterations"]):
            results = detect_sub_matrices(hic_genome, cfg, kernel_matrix, full=full, tsvd=tsvd,
                                          gather_windows=gather_windows)
            coords = [hic_genome.get_full_mat_pattern(d["chr1"], d["chr2"], d["coords"])
                      for d in results if d["coords"] is not None]
            if not coords:
                break  # nothing found with this kernel: next kernel (cli:786-788)
            kshape = kernel_matrix.shape
            wins = [d["windows"] if d["windows"] is not None else np.full((len(d["coords"]),) + kshape, np.nan)
                    for d in results if d["coords"] is not None]
            pile = _global_pileup([d["windows"] for d in results], kshape) if not gather_windows else None
            wins = np.concatenate(wins, axis=0)
            tab = pd.concat(coords, axis=0).reset_index(drop=True)
            tab["kernel_id"] = kernel_id
            tab["iteration"] = it
            all_coords.append(tab)
            all_windows.append(wins)
            kernel_matrix = cud.pileup_patterns(wins) if pile is None else pile  # cli:791
    if not all_coords:
        return None, None
    tab = pd.concat(all_coords, axis=0).reset_index(drop=True)
    wins = np.concatenate(all_windows, axis=0)
    binsize = hic_genome.clr.binsize
    sep = max(int(cfg["min_separation"] // binsize), 1)
    keep = cud.remove_neighbours(tab, win_size=sep)
    tab, wins = tab.loc[keep, :].reset_index(drop=True), wins[keep]
    c1 = hic_genome.bins_to_coords(tab.bin1).reset_index(drop=True)
    c1.columns = [f"{c}1" for c in c1.columns]
    c2 = hic_genome.bins_to_coords(tab.bin2).reset_index(drop=True)
    c2.columns = [f"{c}2" for c in c2.columns]
    tab = pd.concat([tab, c1, c2], axis=1)
    near = (tab.chrom1.astype(str) == tab.chrom2.astype(str)) & \
           (np.abs(tab.start2 - tab.start1) < cfg["min_dist"])
    tab, wins = tab.loc[~near, :], wins[~near.values]
    nanp = tab.pvalue.isnull()
    tab, wins = tab.loc[~nanp, :].reset_index(drop=True), wins[~nanp.values]
    tab["qvalue"] = fdr_correction(tab["pvalue"])
    cols = ["chrom1", "start1", "end1", "chrom2", "start2", "end2", "bin1", "bin2", "kernel_id",
            "iteration", "score", "pvalue", "qvalue"]
    return tab.loc[:, cols], wins


def _locate_positions(positions, hic_genome):
    """Chromosome codes and whole-genome bins of both ends of every position, once for all
    sub-matrices (cli:262-293)."""
    names = [str(c) for c in hic_genome.clr.chromnames]
    from .contacts_map import chrom_codes
    c1 = chrom_codes(positions.chrom1.values, names)
    c2 = chrom_codes(positions.chrom2.values, names)
    b1 = hic_genome.coords_to_bins(pd.DataFrame({"chrom": positions.chrom1.values, "pos": positions.pos1.values}))
    b2 = hic_genome.coords_to_bins(pd.DataFrame({"chrom": positions.chrom2.values, "pos": positions.pos2.values}))
    return names, c1, c2, b1, b2


def _chrom_positions(located, hic_genome, chr1, chr2):
    """Positions falling on one sub-matrix, in sub-matrix bins (cli:262-293): (index into
    the positions, int array [P, 2])."""
    names, c1, c2, b1, b2 = located
    sel = np.flatnonzero((c1 == names.index(str(chr1))) & (c2 == names.index(str(chr2))))
    if len(sel) == 0:
        return sel, np.zeros((0, 2), dtype=np.int64)
    ok = ~(np.isnan(b1[sel]) | np.isnan(b2[sel]))   # positions outside the map are ignored
    sel = sel[ok]
    s1, s2 = hic_genome.clr.extent(chr1)[0], hic_genome.clr.extent(chr2)[0]
    coords = np.stack([b1[sel] - s1, b2[sel] - s2], axis=1).astype(np.int64)
    return sel, coords


def quantify(hic_genome, cfg, bed2d, tsvd=None, return_windows=True, gather_windows=True):
    """Correlation score of every position of `bed2d` (DataFrame chrom1, start1, end1, chrom2,
    start2, end2) with the kernels of `cfg`: the loop of cmd_quantify (cli:295-496), sharded by
    sub-matrix.  `cfg` is modified like the reference does (max_dist = furthest pair,
    min_dist = 0).  Returns (table sorted by bins with score / pvalue / qvalue, windows of the
    best kernel per position or None).  Scores and p-values of all ranks are exchanged as
    float64 record tensors; windows too (gather_windows=True) or they stay on their rank
    (NaN windows for the positions of other ranks)."""
    rank, world, dist = _world()
    bed2d = bed2d.reset_index(drop=True).copy()
    furthest = int(np.max(bed2d.start2 - bed2d.start1))
    cfg["max_dist"] = min(furthest, hic_genome.clr.shape[0] * hic_genome.clr.binsize)
    cfg["min_dist"] = 0
    hic_genome.kernel_config = cfg
    hic_genome.compute_max_dist()
    hic_genome.make_sub_matrices()
    n = len(bed2d)
    positions = bed2d.copy()
    positions["pos1"] = (positions.start1 + positions.end1) // 2
    positions["pos2"] = (positions.start2 + positions.end2) // 2
    km, kn = np.asarray(cfg["kernels"][0]).shape
    mine = set(sharding.partition_units(unit_costs(hic_genome), world)[rank])
    # the output is sorted by bins (cli:476-480): windows are written at their final place
    out_bin1 = hic_genome.coords_to_bins(pd.DataFrame({"chrom": bed2d.chrom1.values, "pos": bed2d.start1.values}))
    out_bin2 = hic_genome.coords_to_bins(pd.DataFrame({"chrom": bed2d.chrom2.values, "pos": bed2d.start2.values}))
    order = np.lexsort((out_bin2, out_bin1))
    rank_of = np.empty(n, dtype=np.int64)
    rank_of[order] = np.arange(n)
    located = _locate_positions(positions, hic_genome)
    best_score = np.full(n, np.nan)
    best_p = np.full(n, np.nan)
    best_win = np.full((n, km, kn), np.nan) if return_windows else None   # in output order
    for kernel_matrix in cfg["kernels"]:
        kernel_matrix = np.asarray(kernel_matrix, dtype=np.float64)
        score = np.full(n, np.nan)
        pval = np.full(n, np.nan)
        wins = {}
        for u, (_, row) in enumerate(hic_genome.sub_mats.iterrows()):
            if u not in mine:
                continue
            idx, coords = _chrom_positions(located, hic_genome, row.chr1, row.chr2)
            if len(idx) == 0:
                continue            # no position on this sub-matrix: not scanned (cli:239-241)
            cm = row.contact_map
            cm.create_mat()
            table, windows = cud.pattern_detector(cm, cfg, kernel_matrix, coords=coords, full=True,
                                                  tsvd=tsvd)
            cm.destroy_mat()
            if table is None:
                continue
            score[idx] = table.score.values
            pval[idx] = table.pvalue.values
            if return_windows:
                wins[u] = (idx, windows)
        if world > 1:
            sel = np.flatnonzero(~np.isnan(score) | ~np.isnan(pval))
            rec = np.column_stack([sel.astype(np.float64), score[sel], pval[sel]])
            for p in sharding.gather_rows(rec):
                ix = p[:, 0].astype(np.int64)
                score[ix], pval[ix] = p[:, 1], p[:, 2]
            if return_windows and gather_windows:
                mine_idx = np.concatenate([i for i, _ in wins.values()]) if wins else np.zeros(0, np.int64)
                mine_w = np.concatenate([w.reshape(len(w), -1) for _, w in wins.values()]) if wins \
                    else np.zeros((0, km * kn))
                idx_parts = sharding.gather_rows(mine_idx.reshape(-1, 1).astype(np.float64))
                win_parts = sharding.gather_rows(np.asarray(mine_w, dtype=np.float64))
                wins = {r: (ip[:, 0].astype(np.int64), wp.reshape(-1, km, kn))
                        for r, (ip, wp) in enumerate(zip(idx_parts, win_parts)) if len(ip)}
        # the best kernel of each position (cli:442-450: highest score; NaN never wins)
        better = (~np.isnan(score)) & (np.isnan(best_score) | (score >= best_score))
        first = np.isnan(best_score) & np.isnan(score) & np.isnan(best_p)
        best_p = np.where(better | first, pval, best_p)
        best_score = np.where(better, score, best_score)
        if return_windows:
            for idx, w in wins.values():
                take = better[idx] | first[idx]
                if take.all():
                    best_win[rank_of[idx]] = w
                else:
                    best_win[rank_of[idx[take]]] = w[take]
    out = bed2d.loc[:, ["chrom1", "start1", "end1", "chrom2", "start2", "end2"]].copy()
    out["bin1"] = out_bin1
    out["bin2"] = out_bin2
    out["score"] = best_score
    out["pvalue"] = best_p
    out["qvalue"] = fdr_correction(out["pvalue"])
    bad = np.isnan(out.score.values)
    out.loc[bad, ["pvalue", "qvalue"]] = np.nan
    out = out.iloc[order].reset_index(drop=True)                # cli:476-480
    return out, best_win
