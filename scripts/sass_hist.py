#!/usr/bin/env python3
"""SASS opcode histogram of every kernel in libchromosight_b200.so (cuobjdump -sass), written to
profiles/<tag>_sass_opcodes.md: the evidence that the build is sm_100a only, that the Pearson
kernel takes its tile through TMA (UTMALDG / UBLKCP + SYNCS mbarrier ops) and runs packed
FFMA2, and which kernels carry float64 (DFMA / DADD) or atomics.

usage: python scripts/sass_hist.py <tag> [path/to/lib.so]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "chromosight_b200", "libchromosight_b200.so")
CUOBJDUMP = os.environ.get("CUOBJDUMP", "/usr/local/cuda/bin/cuobjdump")
sass = subprocess.run([CUOBJDUMP, "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = {}
try:
    names = sorted(set(re.findall(r"Function : (\S+)", sass)))
    out = subprocess.run(["/usr/local/cuda/bin/cu++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    demangle = dict(zip(names, out))
except Exception:
    pass
archs = collections.Counter(re.findall(r"arch = (sm_\w+)", sass))
kernels = collections.OrderedDict()
cur = None
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", line)
    if m and cur is not None:
        cur[m.group(2)] += 1
NOTE = ["UTMALDG", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "DFMA", "DADD", "DMUL", "MUFU",
        "LDS", "STS", "LDG", "STG", "ATOM", "ATOMS", "ATOMG", "RED", "SHFL", "BAR", "LDGSTS"]
path = os.path.join(ROOT, "profiles", f"{tag}_sass_opcodes.md")
with open(path, "w") as f:
    f.write(f"# SASS opcode histogram ({tag})\n\n`cuobjdump -sass {os.path.relpath(lib, ROOT)}` — static instruction counts "
            f"per kernel (not executed counts). Cubins: {dict(archs)}.\n\n")
    f.write("| kernel | SASS instr | " + " | ".join(NOTE) + " |\n|---|---|" + "---|" * len(NOTE) + "\n")
    for k, c in kernels.items():
        base = collections.Counter()
        for op, n in c.items():
            base[op.split(".")[0]] += n
        name = demangle.get(k, k).replace("(int)", "").replace("(bool)", "")
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("cs::", "")
        f.write(f"| `{name}` | {sum(c.values())} | " + " | ".join(str(base.get(o, 0) or "") for o in NOTE) + " |\n")
    f.write("\n## Full opcode lists of the hot kernels\n")
    for k, c in kernels.items():
        name = demangle.get(k, k).replace("(int)", "").replace("(bool)", "")
        if not re.search(r"pearson_tiles<17|emit_rows_narrow|scatter_signal|geo_fill|detrend_rows|exact_windows", name):
            continue
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("cs::", "")
        f.write(f"\n`{name}`: " + ", ".join(f"{op} {n}" for op, n in c.most_common(40)) + "\n")
print("wrote", path, len(kernels), "kernels")
