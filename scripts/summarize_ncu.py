#!/usr/bin/env python3
"""Turn the ncu outputs of scripts/gpu_round.sh into the committed summaries under profiles/:

  profiles/<tag>_launches.md     per-kernel share of one bench step (gpu__time_duration.sum)
  profiles/<tag>_pearson_full.md key metrics, stall reasons and per-section split of the
                                 `ncu --set full` capture of the Pearson kernel
  profiles/pearson_traffic.json  DRAM bytes per launch (read by bench.py for roofline.traffic)

usage: python scripts/summarize_ncu.py <tag>      (reads gpurun_out/*_<tag>.*)
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# ---- launch list
path = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    iN, iV, iG, iB = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    launches = [(r[iN].split("(")[0].replace("void ", ""), float(r[iV]) / 1e3, r[iG], r[iB]) for r in rows[1:]]
    # one bench step = the launches between two consecutive scatter_signal kernels
    starts = [i for i, l in enumerate(launches) if "scatter_signal" in l[0]]
    step = launches[starts[-2]:starts[-1]] if len(starts) >= 2 else launches
    tot = sum(l[1] for l in step)
    agg = collections.OrderedDict()
    for n, t, g, b in step:
        a = agg.setdefault(n, [0, 0.0, g, b])
        a[0] += 1
        a[1] += t
    with open(os.path.join(P, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list, one bench step ({tag})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 3 "
                "--no-cpu-baseline --no-e2e` (cold-cache, serialised launches: compare shares, not absolutes).\n"
                f"{len(launches)} launches captured in total; the table is the last full step "
                f"({len(step)} launches, {tot:.1f} us).\n\n")
        f.write("| kernel | launches | us | share | grid | block |\n|---|---|---|---|---|---|\n")
        for n, (c, t, g, b) in agg.items():
            f.write(f"| `{n}` | {c} | {t:.1f} | {100 * t / tot:.1f}% | {g} | {b} |\n")
    print("wrote", f"profiles/{tag}_launches.md")

# ---- full capture
rep = os.path.join(G, f"prof_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    d = dict(zip(hdr, rows[2]))
    u = dict(zip(hdr, units))
    keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
            'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
            'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
            'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
            'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
            'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.max']
    def to_bytes(v, unit):
        v = float(v)
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    traffic = to_bytes(d['dram__bytes_read.sum'], u['dram__bytes_read.sum']) + \
        to_bytes(d['dram__bytes_write.sum'], u['dram__bytes_write.sum'])
    stalls = []
    for k in hdr:
        if 'average_warps_issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k:
            try:
                stalls.append((float(d[k]), k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')))
            except ValueError:
                pass
    with open(os.path.join(P, f"{tag}_pearson_full.md"), "w") as f:
        f.write(f"# ncu --set full, Pearson kernel ({tag})\n\n")
        f.write("`ncu --set full --clock-control none --import-source on -k regex:pearson -s 3 -c 2 python bench.py "
                "--steps 2 --warmup 3 --no-cpu-baseline --no-e2e` on one B200; first captured launch.\n\n")
        f.write("| metric | value | unit |\n|---|---|---|\n")
        for k in keys:
            if k in d:
                f.write(f"| {k} | {d[k]} | {u.get(k, '')} |\n")
        f.write(f"| dram bytes per launch (read + write) | {traffic / 1e6:.1f} | MB |\n")
        f.write("\nWarp stall reasons (warps per issue slot): " +
                ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]) + "\n")
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                             capture_output=True, text=True).stdout
        tmp = os.path.join("/tmp", f"src_{tag}.csv")
        open(tmp, "w").write(src)
        top = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), tmp, "25"],
                             capture_output=True, text=True).stdout
        f.write("\nHottest source lines (stall samples, share of warp instructions, active threads per instruction):\n\n```\n")
        f.write(top)
        f.write("```\n")
    json.dump({"tag": tag, "kernel": d['Kernel Name'], "dram_bytes_per_launch": traffic,
               "gpu_time_ms": float(d['gpu__time_duration.sum']) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u['gpu__time_duration.sum'], 1)},
              open(os.path.join(P, "pearson_traffic.json"), "w"), indent=1)
    print("wrote", f"profiles/{tag}_pearson_full.md", "traffic MB", traffic / 1e6)
