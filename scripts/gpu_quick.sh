#!/bin/bash
# quick GPU check: parity tests + K1 timing (no CPU legs)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --steps 10 --no-cpu-baseline ${BENCH_ARGS:---no-e2e} 2>gpurun_out/quick.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.3e  ms/step %.3f  breakdown %s  frac %.4f' % (d['value'], d['ms_per_step'], d['step_breakdown_ms'], d['roofline']['frac']))
if d.get('e2e'): print('e2e', d['e2e'])
"
tail -2 gpurun_out/quick.err
