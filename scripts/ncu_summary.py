#!/usr/bin/env python3
"""Summarise an `ncu --page raw --csv` dump: key metrics + stall reasons per kernel.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python scripts/ncu_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.max']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in KEYS:
        if k in d:
            print(f"{k:78s} {d[k]} {rows[1][hdr.index(k)]}")
    st = []
    for k in hdr:
        if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'average_warps_' in k and 'not_issued' not in k:
            try:
                st.append((float(d[k]), k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')))
            except ValueError:
                pass
    print("stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
    print('---')
