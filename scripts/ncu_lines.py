#!/usr/bin/env python3
"""Per-CUDA-source-line totals of an `ncu --page source --csv --print-source cuda,sass` dump
(first profiled launch): stall samples, warp instructions, average active threads.
usage: ncu_lines.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out, fpath, seen_launch = [], None, 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1]
        if fpath.endswith("pearson.cu") or fpath.endswith(".cu"):
            seen_launch += 1
        continue
    if seen_launch > 1:
        break
    if r[0].isdigit():
        try:
            out.append((fpath.split("/")[-1], int(r[0]), r[1].strip(), int(r[6] or 0), int(r[7] or 0), int(r[8] or 0)))
        except ValueError:
            pass
ts = sum(o[3] for o in out); ti = sum(o[4] for o in out)
print(f"total samples {ts}, warp instructions {ti}")
for f, ln, src, s, i, t in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{f[:12]:12s}{ln:5d} {100*s/max(ts,1):5.1f}% smp {100*i/max(ti,1):5.1f}% inst thr/inst {t/max(i,1):5.1f}  {src[:84]}")

# optional: section totals, given "name:lo-hi" arguments after `top`
if len(sys.argv) > 3:
    print("sections:")
    for spec in sys.argv[3:]:
        name, rng = spec.split(":")
        parts = [tuple(int(v) for v in p.split("-")) for p in rng.split(",")]
        s = sum(o[3] for o in out if o[0].endswith(".cu") and any(lo <= o[1] <= hi for lo, hi in parts))
        i = sum(o[4] for o in out if o[0].endswith(".cu") and any(lo <= o[1] <= hi for lo, hi in parts))
        print(f"  {name:14s} {100*s/max(ts,1):5.1f}% samples {100*i/max(ti,1):5.1f}% instructions")
