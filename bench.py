#!/usr/bin/env python3
"""Benchmark of the chromosight hot path on B200 (BASELINE.json metric:
Pearson-windows/s, 17x17 loops kernel, 200k x 200k synthetic intra map).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one sub-matrix, issued exactly as
pattern_detector does (det:253-263): normxcorr2(matrix, kernel, max_dist=D,
sym_upper=True, full=True, missing_mask=mask, missing_tol=0.5, pval=True),
followed by the candidate thresholding of pick_foci (det:417-421) and, for
N > 1, the all-gather of the candidate records.

Legs of the default arm (one JSON line on rank 0):
  value        windows/s with the CSR inputs already resident in HBM (Session.run:
               image fill -> Pearson tiles -> CSR compaction + p-values), CUDA events
               on the launch stream, max over ranks;
  e2e          the same call through chromosight_b200.utils.detection.normxcorr2 with
               host scipy matrices in and out (pinned staging, H2D and D2H inside the
               timed region);
  roofline     the Pearson kernel against the measured HBM peak, 8 B per window;
  cpu_baseline oracle/sparse_port.py (scipy.sparse port of the reference's algorithm)
               on a bounded row-slab of the same map, 1 core.
`--impl reference` times that port on all host cores (one slab per worker per step).
Windows are counted by the metric's definition, sum_{d<=D}(n-d), although the
reference call (and therefore this one) also evaluates the k+(k-1) diagonals
beyond D that pattern_detector trims afterwards (det:270).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "pearson_windows_per_s"
UNIT = "windows/s"
BYTES_PER_WINDOW = 8  # one fp32 pixel in + one fp32 score out


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=200_000, help="bins of the synthetic chromosome")
    ap.add_argument("--max-dist", type=int, default=200, help="scan distance in bins (2 Mb @ 10 kb)")
    ap.add_argument("--kernel", default="loops")
    ap.add_argument("--win-size", type=int, default=0,
                    help="resize the kernel to this (odd) width like `--win-size` of the CLI (cli:690-695)")
    ap.add_argument("--kernel-index", type=int, default=0)
    ap.add_argument("--pearson", type=float, default=0.3)
    ap.add_argument("--cpu-rows", type=int, default=50_000, help="rows of the CPU-baseline slab")
    ap.add_argument("--ref-rows", type=int, default=2_000, help="rows per worker per step (--impl reference)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 8)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: one chromosome per GPU (independent sub-matrices); strong: ONE chromosome cut "
                         "into row slabs over the GPUs (rowslab.py: law all-reduce + candidate gather)")
    ap.add_argument("--config", type=int, default=0, choices=[0, 4, 5],
                    help="4: quantify loops at --positions bed2d positions on 4 chromosomes x 50k bins, sharded "
                         "by chromosome; 5: detect loops --inter on a synthetic 23-chromosome / 500k-bin genome, "
                         "sharded by sub-matrix (BASELINE.json configs[3], [4]); 0: the metric workload")
    ap.add_argument("--positions", type=int, default=1_000_000)
    ap.add_argument("--genome-bins", type=int, default=500_000, help="config 5: bins of the whole genome")
    ap.add_argument("--chroms", type=int, default=23, help="config 5: chromosomes (sizes ~ hg38)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------- workload
def raw_map(n, D, k, seed):
    from chromosight_b200 import synthetic
    return synthetic.band_counts(n, D + k, seed=seed, missing_frac=0.02, max_dist=D)


def finish_map(mat, detect, D, k, trim, mask_fn):
    """What ContactMap.create_mat does after detrending (cm:618-624, cm:539-548) and the
    mask pattern_detector builds (det:242-248)."""
    mat = trim(mat.tocsr(), D + k)
    mat.data[np.isnan(mat.data)] = 0
    mat.eliminate_zeros()
    mask = mask_fn(mat.shape, detect, detect, max_dist=D, sym_upper=True)
    return mat, mask


def call_kwargs(D):
    return dict(max_dist=D, sym_upper=True, full=True, missing_tol=0.5, pval=True)


# --------------------------------------------------------------------------- CPU arms
def _cpu_impl():
    """The CPU implementation that is timed: the UNMODIFIED reference (oracle/_ref, copied from
    /root/reference by oracle/make_ref.sh; kind "reference") when it travelled with the
    repository, else the scipy.sparse port of its algorithm (oracle/sparse_port.py; kind "port",
    measured 1.3-1.6x faster than the reference)."""
    from oracle import ref_loader
    ref = ref_loader.load()
    if ref is not None:
        det, pre, _ = ref
        return ("reference", lambda raw, detect, md: pre.detrend(raw, detectable_bins=detect, max_dist=md, max_val=10),
                det.normxcorr2, pre.diag_trim, pre.make_missing_mask)
    from chromosight_b200.utils import preprocessing as hostpre  # host-only helpers (no CUDA)
    from oracle import sparse_port as spt
    return ("port", lambda raw, detect, md: spt.detrend_sparse(raw, detect, md, 10),
            spt.normxcorr2_sparse, hostpre.diag_trim, hostpre.make_missing_mask)


def _cpu_slab(args):
    """One CPU unit of work: raw slab-map -> detrend (pre:256-310) -> Pearson map (det:807-914),
    both timed, with the reference's own code where available."""
    rows, D, kernel, seed, reps = args
    import warnings
    warnings.simplefilter("ignore")
    kind, detrend, normxcorr2, trim, mask_fn = _cpu_impl()
    k = kernel.shape[0]
    raw, detect = raw_map(rows, D, k, seed)
    raw = raw.tocsr()
    t0 = time.perf_counter()
    mat = detrend(raw, detect, D + k)
    dt_detrend = time.perf_counter() - t0
    nnz_raw = int(raw.nnz)
    mat, mask = finish_map(mat, detect, D, k, trim, mask_fn)
    t0 = time.perf_counter()
    for _ in range(reps):
        r, p = normxcorr2(mat, kernel, missing_mask=mask, **call_kwargs(D))
    dt = (time.perf_counter() - t0) / reps
    return dt, int(r.nnz), kind, dt_detrend, nnz_raw


def cpu_baseline(rows, D, kernel):
    from chromosight_b200 import synthetic
    dt, _, kind, dt_detrend, nnz_raw = _cpu_slab((rows, D, kernel, 0, 1))
    nwin = synthetic.n_windows(rows, D)
    what = ("chromosight.utils.detection.normxcorr2 of the unmodified reference (oracle/_ref)" if kind == "reference"
            else "oracle/sparse_port.normxcorr2_sparse (scipy.sparse port of det:917-1131)")
    return {"value": nwin / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"one slab of the same generator, as the reference processes a chromosome (one core per "
                      f"sub-matrix, cli:738-755): {rows} rows x D={D}, {nwin} windows, {dt:.2f} s, {what}; "
                      f"host has {os.cpu_count()} cpus",
            "detrend": {"value": nnz_raw / dt_detrend, "unit": "nnz/s", "seconds": dt_detrend, "nnz": nnz_raw,
                        "what": "preprocessing.detrend (pre:256-310) of the same slab, 1 core"}}


def run_reference(a, kernel):
    """--impl reference: the reference's own CPU implementation of the path on every host core;
    a step = one slab of `ref_rows` rows per worker (the reference parallelises over
    sub-matrices with one core each, cli:738-755)."""
    import multiprocessing as mp
    from chromosight_b200 import synthetic
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    D, rows = a.max_dist, a.ref_rows
    nwin_slab = synthetic.n_windows(rows, D)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        jobs = [(rows, D, kernel, 1000 + i, 1) for i in range(cores)]
        for _ in range(a.warmup):
            pool.map(_cpu_slab, jobs)
        t0 = time.perf_counter()
        per_step = []
        kind = "port"
        for _ in range(a.steps):
            # input generation runs in the workers but outside the timer: a step costs the
            # slowest worker's normxcorr2 time
            res = pool.map(_cpu_slab, jobs)
            per_step.append(max(r[0] for r in res))
            kind = res[0][2]
        wall = time.perf_counter() - t0
    t = float(np.sum(per_step))
    value = cores * nwin_slab * a.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(a, kernel, sample=f"{cores} slabs of {rows} rows per step"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{cores} workers x {rows}-row slab ({nwin_slab} windows each) per "
                                   f"step, max-over-workers time per step, wall {wall:.1f} s incl. input generation"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(a, kernel, sample=None):
    from chromosight_b200 import synthetic
    k = kernel.shape[0]
    cfg = {
        "workload": f"intra {a.n}x{a.n} synthetic band map (10 kb bins), max_dist {a.max_dist} bins, "
                    f"{a.kernel} kernel {k}x{k}, masked full-mode normxcorr2 as issued by pattern_detector",
        "n_bins": a.n, "max_dist_bins": a.max_dist, "kernel": f"{a.kernel} {k}x{k}",
        "windows_per_map": synthetic.n_windows(a.n, a.max_dist),
        "missing_bins": "2%", "seed": 0,
        "l2": "inputs_exceed_l2 (CSR input, fp32 band and score band are each > 126 MB)",
        "parallelism": (f"one chromosome cut into {a.gpus} row slabs (halo k above, D+3k below; distance law by one all-reduce)"
                        if getattr(a, "scaling", "weak") == "strong" and a.gpus > 1
                        else f"{a.gpus} x one chromosome per GPU (independent sub-matrices)"),
    }
    if sample:
        cfg["sample"] = sample
    return cfg


# --------------------------------------------------------------------------- GPU arm
class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.f = None

    def start(self):
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def run_b200(a, kernel):
    import torch
    import torch.distributed as dist
    from chromosight_b200 import _lib, sharding, synthetic
    from chromosight_b200.session import Session
    from chromosight_b200.utils import detection as cud, preprocessing as cup

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the hot path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the one JSON line (NCCL prints its version banner there otherwise)
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    n, D, k = a.n, a.max_dist, kernel.shape[0]
    kw = call_kwargs(D)
    nwin = synthetic.n_windows(n, D)

    strong = a.scaling == "strong" and world > 1
    plan = None
    if strong:
        # ---- ONE chromosome over all ranks: row slabs with halos, global distance law by one
        # all-reduce of the per-diagonal sums
        from chromosight_b200 import rowslab
        raw, detect_all = raw_map(n, D, k, seed=0)
        plan = rowslab.slab_plan(n, world, k, D)[rank]
        law = rowslab.global_law(raw, detect_all, D + k, plan[0], plan[1])
        mat, detect = rowslab.slab_inputs(raw, detect_all, law, plan, D, k)
        mask = cup.make_missing_mask(mat.shape, detect, detect, max_dist=D, sym_upper=True)
    else:
        # ---- one chromosome per rank (weak scaling), detrended with the CUDA path (a8)
        raw, detect = raw_map(n, D, k, seed=rank)
        mat = cup.detrend(raw, detectable_bins=detect, max_dist=D + k, max_val=10)
        mat, mask = finish_map(mat, detect, D, k, cup.diag_trim, cup.make_missing_mask)
    detrend_leg = detrend_device_leg(raw, detect_all if strong else detect, D + k, torch) if rank == 0 else None
    del raw

    sess = Session(local)
    # (strong scaling: a rank scores the rows it owns, not the halo around them)
    sess.upload(mat, kernel, missing_mask=mask,
                out_rows=(plan[0] - plan[2], plan[1] - plan[2]) if strong else None, **kw)
    cap = 1 << 22
    cand_buf = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    state = {"ncand": 0, "gathered": 0}
    gather_cap = 1 << 16   # records exchanged per rank by the final gather (1 MB)

    def step():
        # the run is enqueued, the candidate thresholding right behind it: ONE host
        # synchronisation per step (the run's checks are made there)
        sess.run(wait=False)
        _, nc = sess.candidates(a.pearson, 0, D, out=cand_buf)
        state["ncand"] = nc
        return sess.wait()

    def final_gather():
        # the one collective of the path (north_star): the candidate records of every rank, once,
        # at the end of the run; fixed size, no host synchronisation
        if world > 1:
            send, nc = cand_buf, state["ncand"]
            if strong:
                from chromosight_b200 import rowslab
                from chromosight_b200.session import records_to_numpy
                mine = rowslab.owned_candidates(records_to_numpy(cand_buf, nc), plan)
                nc = len(mine)
                send = torch.zeros((gather_cap, 4), dtype=torch.int32, device=dev)
                send[:nc] = torch.from_numpy(mine.view(np.int32).reshape(-1, 4)).to(dev)
            gathered, counts = sharding.gather_candidates(send, nc, cap=gather_cap)
            state["gathered_t"] = (gathered, counts)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        st = step()
    # the collective is warmed up like the kernels (NCCL sets its channels up on the first call of a
    # kind: ~5 ms at 8 ranks, which belongs to no step)
    final_gather()
    launches_per_step = None
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms_pearson, ms_fill, ms_compact = [], [], []
    ev0.record()
    for _ in range(a.steps):
        st = step()
        ms_pearson.append(st["ms_pearson"])
        ms_fill.append(st["ms_fill"])
        ms_compact.append(st["ms_compact"])
    ev_loop = torch.cuda.Event(enable_timing=True)
    ev_loop.record()          # this rank's own steps, before it meets the others in the gather
    final_gather()
    ev1.record()
    barrier()
    if world > 1:
        state["gathered"] = int(state["gathered_t"][1].sum().item())
    launches = _lib.launch_count() - l0
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ev0.elapsed_time(ev_loop)], dtype=torch.float64, device=dev)
    per_rank_ms = [float(t.item()) / a.steps]
    if world > 1:
        # every rank's own device time (the spread between GPUs), then the max the value is quoted on
        allt = torch.zeros(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allt, t)
        per_rank_ms = [float(x) / a.steps for x in allt.tolist()]
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    else:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    ms_step = float(t.item()) / a.steps
    value = (1 if strong else world) * nwin / (ms_step * 1e-3)
    n_eval = int(st["n_windows"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    t_k = float(np.mean(ms_pearson)) * 1e-3
    achieved = BYTES_PER_WINDOW * n_eval / t_k / 1e9
    roofline = {
        "bound": "hbm", "kernel": "pearson_tiles", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak,
        "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
        "traffic": traffic_from_profile(),
        "bytes_per_window": BYTES_PER_WINDOW, "windows_per_launch": n_eval,
        "kernel_ms": t_k * 1e3,
        "fp32_fma_per_window": int(kernel.size),
        "note": "CUDA-core stencil: k*k FMAs per window bound it well below the HBM roof (DESIGN.md)",
    }

    # ---- e2e: the reference-facing call, host matrices in and out
    e2e = None
    detector_e2e = None
    if not a.no_e2e:
        ke = a.e2e_steps or min(a.steps, 8)
        from chromosight_b200 import _cuda
        # the contract's input side: the step's inputs sit in pinned host memory (the same
        # scipy matrices, their arrays page-locked); the pageable variant is timed as well

        def timed(m, mk):
            for _ in range(2):
                r, p = cud.normxcorr2(m, kernel, missing_mask=mk, **kw)
            del r, p
            barrier()
            t0 = time.perf_counter()
            for _ in range(ke):
                r, p = cud.normxcorr2(m, kernel, missing_mask=mk, **kw)
                cs = float(r.data[:: max(1, r.nnz // 1024)].sum())  # touch the result on the host
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / ke, r, p, cs

        # the drop-in case leads: ordinary (pageable) scipy matrices, staged through the
        # library's pinned buffers; page-locked inputs (direct DMA) are timed next to it
        te, r, p, checksum = timed(mat, mask)
        s = dict(cud.last_call_stats)
        nnz_e2e = int(r.nnz)
        del r, p
        te_pinned, r, p, _ = timed(_cuda.pin_sparse(mat), _cuda.pin_sparse(mask))
        del r, p
        t = torch.tensor([te, te_pinned], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        te, te_pinned = float(t[0].item()), float(t[1].item())
        e2e = {"value": (1 if strong else world) * nwin / te, "unit": UNIT,
               "h2d_bytes_per_step": int(s.get("h2d_bytes", 0)), "d2h_bytes_per_step": int(s.get("d2h_bytes", 0)),
               "ms_per_step": te * 1e3, "steps": ke, "timer": "host wall clock around the API call, max over ranks",
               "inputs": "pageable scipy CSR matrices (signal float64 + missing mask), staged through pinned buffers",
               "outputs": "two scipy CSR float64 matrices (scores, log10 p) of every non-zero score, as the reference returns",
               "wire_format": "float32 score + float32 log10p + uint8 diagonal offset per stored score, widened to "
                              "float64 / int32 on the host (csrc/host_expand.cpp)",
               "pinned_inputs_ms_per_step": te_pinned * 1e3,
               "overlapped_span_ms": s.get("ms_kernels"),
               "result_nnz": nnz_e2e, "checksum": checksum}

        # ---- detector e2e: the reference's real call site (pattern_detector, det:177-345):
        # host matrix in, table of loops + their windows out; only foci leave the device
        from chromosight_b200 import kernels as presets
        cfg = dict(getattr(presets, a.kernel))
        cfg["pearson"] = a.pearson
        # the synthetic generator thins the map out with distance (half the pixels are zeros at
        # 0.5 Mb): keep the zero-pixel filter of validate_patterns from discarding every loop
        cfg["max_perc_zero"] = 100.0

        class _Map:  # the attributes pattern_detector reads of a ContactMap (det:230-257)
            matrix, detectable_bins, max_dist, inter, name = mat, (detect, detect), D, False, "bench"

        def detect_once():
            return cud.pattern_detector(_Map, cfg, kernel, full=True)

        for _ in range(2):
            tab, wins = detect_once()
        barrier()
        kd = max(3, min(ke, 5))
        t0 = time.perf_counter()
        for _ in range(kd):
            tab, wins = detect_once()
        td = (time.perf_counter() - t0) / kd
        t = torch.tensor([td], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        td = float(t.item())
        detector_e2e = {"value": (1 if strong else world) * nwin / td, "unit": UNIT, "ms_per_step": td * 1e3, "steps": kd,
                        "call": "chromosight_b200.utils.detection.pattern_detector(contact_map, config, kernel, full=True)",
                        "patterns": 0 if tab is None else int(len(tab)),
                        "h2d_bytes_per_step": int(s.get("h2d_bytes", 0)),
                        "d2h": "foci records, k x k float64 windows and the score / p-value at the foci"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
        "dtype_detail": "float32 image, sums and score (centred algebra), float64 exact path for ill-conditioned / "
                        "near-threshold windows; CSR results float64",
        "data": "synthetic", "config": workload_config(a, kernel),
        "roofline": roofline, "e2e": e2e, "detector_e2e": detector_e2e, "detrend": detrend_leg,
        "gpu_launches": int(launches), "clocks": clocks,
        "step_breakdown_ms": {"fill": float(np.mean(ms_fill)), "pearson": float(np.mean(ms_pearson)),
                              "compact_csr_pvalues": float(np.mean(ms_compact))},
        "per_rank_ms_per_step": per_rank_ms,
        "candidates_per_map": state["ncand"], "candidates_gathered": state["gathered"],
        "collective": "one all-gather of the candidate records (fixed 65536 x 16 B per rank) after the last step"
                      if world > 1 else None,
        "result_nnz": int(st["nnz"]),
    }
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(a.cpu_rows, D, kernel)
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def detrend_device_leg(raw, detect, max_dist, torch, reps=10):
    """K0a (csrc/detrend.cu: distance law + division, pre:129-310) with the raw CSR resident in
    HBM: CUDA events around the two library calls, algorithmic bytes 20 B per stored pixel read
    (8 B value + 4 B column twice: law pass and division pass... counted once each) + 8 B written."""
    import ctypes as C
    from chromosight_b200 import _cuda, _lib
    lib = _lib.load()
    csr = raw.tocsr()
    n = csr.shape[0]
    n_diags = int(min(n, max_dist + 1))
    flags = np.zeros(n, dtype=np.uint8)
    flags[np.asarray(detect)] = 1
    d = dict(indptr=_cuda.to_device(csr.indptr, np.int64), indices=_cuda.to_device(csr.indices, np.int32),
             data=_cuda.to_device(csr.data, np.float64), det=_cuda.to_device(flags))
    d_sum, d_cnt = _cuda.empty(n_diags, torch.float64), _cuda.empty(n_diags, torch.int64)
    d_law, d_out = _cuda.empty(n, torch.float64), _cuda.empty(csr.nnz, torch.float64)

    def once():
        _lib.check(lib.cs_distance_law(_cuda.ptr(d["indptr"]), _cuda.ptr(d["indices"]), _cuda.ptr(d["data"]), n,
                                       _cuda.ptr(d["det"]), n_diags, _cuda.ptr(d_sum), _cuda.ptr(d_cnt),
                                       _cuda.ptr(d_law), _cuda.stream_ptr()))
        _lib.check(lib.cs_detrend_apply(_cuda.ptr(d["indptr"]), _cuda.ptr(d["indices"]), _cuda.ptr(d["data"]),
                                        _cuda.ptr(d_out), n, _cuda.ptr(d_law), n, C.c_double(10.0),
                                        _cuda.stream_ptr()))

    for _ in range(3):
        once()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        once()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nnz = int(csr.nnz)
    algo = 32 * nnz   # law pass reads value + column (12 B), division pass reads 12 B and writes 8 B
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    return {"kernel": "diag_accumulate + law_finalize + detrend_rows (csrc/detrend.cu)", "ms": ms, "nnz": nnz,
            "value": nnz / (ms * 1e-3), "unit": "stored pixels/s",
            "roofline": {"bound": "hbm", "achieved": algo / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": algo / (ms * 1e-3) / 1e9 / peak, "bytes_per_nnz": 32,
                         "note": "two passes over the CSR entries (law: 12 B read; division: 12 B read + 8 B written)"}}


# --------------------------------------------------------------------------- configs 4 and 5
HG38_MB = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156]


def _dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the hot path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
    return world, rank, dist, sync


def run_config4(a):
    """BASELINE.json configs[3]: quantify loops at 1 M bed2d positions on the 200k map cut into 4
    chromosomes of 50k bins, sharded by chromosome (driver.quantify = cmd_quantify, cli:295-496)."""
    import pandas as pd
    from chromosight_b200 import driver, kernels, synthetic
    from chromosight_b200.contacts_map import HicGenome
    from chromosight_b200.utils import detection as cud
    world, rank, dist, sync = _dist_setup()
    binsize, D, nb = 10_000, a.max_dist, a.n // 4
    t0 = time.perf_counter()
    clr = synthetic.genome_cool([nb] * 4, binsize=binsize, n_diags=D + 17, seed=0, density_floor=1.0)
    t_gen = time.perf_counter() - t0
    rng = np.random.default_rng(1)
    P = a.positions
    chrom = rng.integers(0, 4, size=P)
    b1 = rng.integers(0, nb - D, size=P)
    b2 = b1 + rng.integers(2, D + 1, size=P)
    names = np.array(clr.chromnames)[chrom]
    bed = pd.DataFrame({"chrom1": names, "start1": b1 * binsize, "end1": (b1 + 1) * binsize,
                        "chrom2": names, "start2": b2 * binsize, "end2": (b2 + 1) * binsize})
    cfg = dict(kernels.loops)
    cfg["kernels"] = [np.array(k) for k in cfg["kernels"]]
    hg = HicGenome(clr, inter=False, kernel_config=cfg)
    hg.normalize()
    driver.quantify(hg, cfg, bed.iloc[:2000].copy(), return_windows=False)   # warm-up
    for key in cud.detector_totals:
        cud.detector_totals[key] = 0
    sync()
    t0 = time.perf_counter()
    table, windows = driver.quantify(hg, cfg, bed, return_windows=True, gather_windows=False)
    sync()
    dt = time.perf_counter() - t0
    if rank == 0:
        tot = cud.detector_totals
        print(json.dumps({
            "config": 4, "what": "quantify loops (driver.quantify, cli:295-496)", "positions": P, "chroms": 4,
            "bins_per_chrom": nb, "n_gpus": world, "seconds": dt, "positions_per_s": P / dt,
            "scored": int(table.score.notna().sum()),
            "windows": "k x k float64 window per position, kept on the rank that cut it",
            "rank0_pattern_detector": {"calls": tot["calls"], "wall_s": tot["wall_ms"] / 1e3,
                                       "device_s": tot["device_ms"] / 1e3},
            "host_share": "the rest of `seconds`: sub-matrix extraction and balancing from the pixel table, "
                          "detrend round trip, pandas bookkeeping of cmd_quantify",
            "reference_estimate": "validate_patterns alone: 207 us per position (SURVEY 6) = %.0f s" % (P * 207e-6),
            "genome_generation_s": t_gen}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_config5(a):
    """BASELINE.json configs[4]: detect loops --inter on a synthetic 23-chromosome genome of 500k
    bins (sizes proportional to hg38), 23 intra + 253 inter sub-matrices sharded over the ranks
    by window count (driver.detect = cmd_detect, cli:625-860)."""
    from chromosight_b200 import driver, kernels, synthetic
    from chromosight_b200.contacts_map import HicGenome
    from chromosight_b200.utils import detection as cud
    world, rank, dist, sync = _dist_setup()
    w = np.array(HG38_MB[: a.chroms], dtype=np.float64)
    sizes = np.maximum((w / w.sum() * a.genome_bins).astype(int), 200)
    t0 = time.perf_counter()
    clr = synthetic.genome_cool([int(x) for x in sizes], binsize=10_000, n_diags=a.max_dist + 17, seed=10,
                                inter_density=1e-4, density_floor=1.0)
    t_gen = time.perf_counter() - t0
    cfg = dict(kernels.loops)
    cfg["kernels"] = [np.array(k) for k in cfg["kernels"]]
    hg = HicGenome(clr, inter=True, kernel_config=cfg)
    hg.normalize()
    hg.make_sub_matrices()
    costs = driver.unit_costs(hg)
    for key in cud.detector_totals:
        cud.detector_totals[key] = 0
    sync()
    t0 = time.perf_counter()
    table, windows = driver.detect(hg, cfg, full=True, gather_windows=False)
    sync()
    dt = time.perf_counter() - t0
    if rank == 0:
        tot = cud.detector_totals
        n_inter = 0 if table is None else int((table.chrom1.astype(str) != table.chrom2.astype(str)).sum())
        print(json.dumps({
            "config": 5, "what": "detect loops --inter (driver.detect, cli:625-860)", "chroms": int(a.chroms),
            "genome_bins": int(sizes.sum()), "sub_matrices": len(costs), "windows_total": float(np.sum(costs)),
            "n_gpus": world, "seconds": dt, "windows_per_s": float(np.sum(costs)) / dt,
            "patterns": 0 if table is None else len(table), "inter_patterns": n_inter,
            "rank0_pattern_detector": {"calls": tot["calls"], "wall_s": tot["wall_ms"] / 1e3,
                                       "device_s": tot["device_ms"] / 1e3},
            "host_share": "the rest of `seconds`: sub-matrix extraction from the pixel table, inter "
                          "normalisation, table bookkeeping of cmd_detect",
            "genome_generation_s": t_gen}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def traffic_from_profile():
    """DRAM bytes per launch of the Pearson kernel from the committed ncu capture
    (profiles/pearson_traffic.json, written by scripts/summarize_ncu.py), else null."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "pearson_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        return None


def main():
    a = parse_args()
    from chromosight_b200 import kernels
    kernel = np.asarray(getattr(kernels, a.kernel)["kernels"][a.kernel_index], dtype=np.float64)
    if a.win_size:
        from chromosight_b200.utils import preprocessing as hostpre
        kernel = hostpre.resize_kernel(kernel, factor=a.win_size / kernel.shape[0])
    if a.impl == "reference":
        run_reference(a, kernel)
    elif a.config == 4:
        run_config4(a)
    elif a.config == 5:
        run_config5(a)
    else:
        run_b200(a, kernel)


if __name__ == "__main__":
    main()
