#!/usr/bin/env python3
"""BASELINE.json configs 4 and 5 at (scaled) size on the GPU box, through the sharded drivers.

  config 4: quantify loops at P bed2d positions on a synthetic 4-chromosome genome
            (4 x 50 000 bins = the 200k map cut in four), sharded by chromosome;
  config 5: detect loops --inter on a synthetic multi-chromosome genome, sharded by sub-matrix
            (scaled: C chromosomes of ~B bins instead of 23 chromosomes / 500k bins, because
            inter maps are dense: ms x ns float32 images).

Launch with python (1 GPU) or torchrun (N GPUs).  Prints one JSON line per config on rank 0.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pandas as pd
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--positions", type=int, default=1_000_000)
ap.add_argument("--chrom-bins", type=int, default=50_000)
ap.add_argument("--inter-chroms", type=int, default=4)
ap.add_argument("--inter-bins", type=int, default=6000)
ap.add_argument("--skip4", action="store_true")
ap.add_argument("--skip5", action="store_true")
a = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

from chromosight_b200 import driver, kernels, synthetic
from chromosight_b200.contacts_map import HicGenome


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


if not a.skip4:
    binsize, D = 10_000, 200
    t0 = time.perf_counter()
    clr = synthetic.genome_cool([a.chrom_bins] * 4, binsize=binsize, n_diags=D + 17, seed=0, density_floor=1.0)
    t_gen = time.perf_counter() - t0
    rng = np.random.default_rng(1)
    P = a.positions
    chrom = rng.integers(0, 4, size=P)
    b1 = rng.integers(0, a.chrom_bins - D, size=P)
    b2 = b1 + rng.integers(2, D + 1, size=P)
    names = np.array(clr.chromnames)[chrom]
    bed = pd.DataFrame({"chrom1": names, "start1": b1 * binsize, "end1": (b1 + 1) * binsize,
                        "chrom2": names, "start2": b2 * binsize, "end2": (b2 + 1) * binsize})
    cfg = dict(kernels.loops); cfg["kernels"] = [np.array(k) for k in cfg["kernels"]]
    hg = HicGenome(clr, inter=False, kernel_config=cfg); hg.normalize()
    # warm-up on a small table (CUDA context, buffers)
    driver.quantify(hg, cfg, bed.iloc[:2000].copy(), return_windows=False)
    sync(); t0 = time.perf_counter()
    if os.environ.get("PROFILE") and rank == 0:
        import cProfile, pstats, io
        pr = cProfile.Profile(); pr.enable()
    table, windows = driver.quantify(hg, cfg, bed, return_windows=(world == 1))
    if os.environ.get("PROFILE") and rank == 0:
        pr.disable(); s_ = io.StringIO(); pstats.Stats(pr, stream=s_).sort_stats("cumulative").print_stats(30)
        sys.stderr.write(s_.getvalue()[:7000])
    sync(); dt = time.perf_counter() - t0
    if rank == 0:
        ok = int(table.score.notna().sum())
        print(json.dumps({"config": 4, "what": "quantify loops", "positions": P, "chroms": 4,
                          "bins_per_chrom": a.chrom_bins, "n_gpus": world, "seconds": dt,
                          "positions_per_s": P / dt, "scored": ok,
                          "windows_returned": None if windows is None else list(windows.shape),
                          "reference_estimate_s": "207 us per coordinate in validate_patterns alone (SURVEY 6) -> %.0f s" % (P * 207e-6),
                          "genome_generation_s": t_gen}), flush=True)
    del table, windows, bed, hg, clr

if not a.skip5:
    binsize = 10_000
    rng = np.random.default_rng(2)
    sizes = [int(a.inter_bins * f) for f in np.linspace(1.0, 0.6, a.inter_chroms)]
    clr = synthetic.genome_cool(sizes, binsize=binsize, n_diags=217, seed=10, inter_density=2e-3, density_floor=1.0)
    cfg = dict(kernels.loops); cfg["kernels"] = [np.array(k) for k in cfg["kernels"]]
    hg = HicGenome(clr, inter=True, kernel_config=cfg); hg.normalize(); hg.make_sub_matrices()
    costs = driver.unit_costs(hg)
    sync(); t0 = time.perf_counter()
    table, windows = driver.detect(hg, cfg, full=True)
    sync(); dt = time.perf_counter() - t0
    if rank == 0:
        n_inter = int((table.chrom1.astype(str) != table.chrom2.astype(str)).sum()) if table is not None else 0
        print(json.dumps({"config": 5, "what": "detect loops --inter", "chrom_bins": sizes,
                          "sub_matrices": len(costs), "windows_total": float(np.sum(costs)),
                          "n_gpus": world, "seconds": dt, "windows_per_s": float(np.sum(costs)) / dt,
                          "patterns": 0 if table is None else len(table), "inter_patterns": n_inter}), flush=True)
if world > 1:
    dist.destroy_process_group()
