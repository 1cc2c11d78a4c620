#!/bin/bash
# marginal cost of the Pearson kernel's phases: timing with parts switched off (results are wrong
# on purpose; CS_DEBUG_SKIP bits: 1 mask sums, 2 score formulas, 4 sums pass, 8 FMA pass, 32 redo)
for m in ${MASKS:-0 47 63 111 175 303 319 447 511}; do
  echo -n "skip=$m: "
  CS_DEBUG_SKIP=$m timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pearson ms %.3f' % d['step_breakdown_ms']['pearson'])"
done
