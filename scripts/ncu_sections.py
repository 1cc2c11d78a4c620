#!/usr/bin/env python3
"""Per-section totals of a Pearson-kernel `ncu --page source --csv --print-source cuda,sass` dump:
sections are delimited by marker comments `// [sec:NAME]` in pearson.cu (a section runs to the
next marker).  usage: ncu_sections.py src.csv [path/to/pearson.cu] [windows]"""
import csv, collections, re, sys
src_csv = sys.argv[1]
cu = sys.argv[2] if len(sys.argv) > 2 else "chromosight_b200/csrc/pearson.cu"
nwin = float(sys.argv[3]) if len(sys.argv) > 3 else 46772739.0
marks = []
for i, l in enumerate(open(cu), 1):
    m = re.search(r"\[sec:(\w+)\]", l)
    if m: marks.append((i, m.group(1)))
def sec(ln):
    name = "pre"
    for i, n in marks:
        if ln >= i: name = n
    return name
rows = list(csv.reader(open(src_csv)))
launch = 0; cur = None; fpath = ""
inst = collections.Counter(); thr = collections.Counter(); smp = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for r in rows:
    if not r: continue
    if r[0] == "File Path":
        fpath = r[1]; launch += fpath.endswith("pearson.cu"); continue
    if launch > 1: break
    if r[0].isdigit():
        cur = sec(int(r[0])) if fpath.endswith("pearson.cu") else "inl:" + fpath.split("/")[-1][:18]; continue
    if len(r) > 8 and r[0] == "" and r[2].startswith("0x"):
        n = int(r[7] or 0); t = int(r[8] or 0); s = int(r[6] or 0)
        inst[cur] += n; thr[cur] += t; smp[cur] += s
        p = r[3].split(); op = p[1] if p[0].startswith("@") else p[0]
        ops[cur][op.split(".")[0]] += n
ti = sum(inst.values()); ts = sum(smp.values())
print(f"warp instructions {ti/1e6:.1f}M, samples {ts}")
print(f"{'section':22s} {'Minst':>8s} {'%inst':>6s} {'%smp':>6s} {'thr-inst/win':>12s}  top opcodes (M)")
for k, n in sorted(inst.items(), key=lambda kv: -smp[kv[0]]):
    print(f"{k:22s} {n/1e6:8.1f} {100*n/ti:6.1f} {100*smp[k]/ts:6.1f} {thr[k]/nwin:12.1f}  " + ", ".join(f"{o}:{c/1e6:.0f}" for o, c in ops[k].most_common(6)))
