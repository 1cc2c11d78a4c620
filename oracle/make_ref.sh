#!/bin/bash
# Recipe for oracle/_ref: the UNMODIFIED reference implementation of the hot path
# (koszullab/chromosight, pure Python), taken from the sources where they lie under
# /root/reference.  oracle/_ref/ is git-ignored (reference sources never enter the history)
# but travels to the GPU box with the repository snapshot, like the built .so files, so that
#   * bench.py --impl reference and bench.py's cpu_baseline time the reference itself
#     (cpu_baseline.kind = "reference"),
#   * tests can validate the oracle against it where it is present.
# Only the modules of the path are taken: utils/detection.py (normxcorr2, xcorr2, pick_foci ...),
# utils/preprocessing.py (detrend, masks), utils/stats.py (corr_to_pval).
set -e
SRC=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
DST=$HERE/_ref
[ -d "$SRC/chromosight" ] || { echo "no reference tree at $SRC" >&2; exit 1; }
rm -rf "$DST"
mkdir -p "$DST/chromosight/utils"
for f in __init__.py version.py utils/__init__.py utils/detection.py utils/preprocessing.py utils/stats.py; do
  cp "$SRC/chromosight/$f" "$DST/chromosight/$f"
done
(cd "$SRC" && git rev-parse HEAD 2>/dev/null || echo unknown) > "$DST/COMMIT"
echo "oracle/_ref: $(ls "$DST/chromosight" "$DST/chromosight/utils" | wc -l) entries from $SRC"
