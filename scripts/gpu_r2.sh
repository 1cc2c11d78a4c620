#!/bin/bash
# One GPU-box session of round 2.  Usage:
#   gpurun --timeout 1500 -- 'bash scripts/gpu_r2.sh <tag> [what...]'
# what: tests | smoke | bench | benchq (kernel timing only) | ablate | count | ncu | sass
TAG=${1:-r2}; shift
WHAT=${@:-tests benchq}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
quick() {  # kernel timing of one bench configuration: $1 label, rest = bench args
  local label=$1; shift
  timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-e2e "$@" 2>gpurun_out/quick_$TAG.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$label value %.3e  ms/step %.3f  breakdown %s  frac %.4f' % (d['value'], d['ms_per_step'], d['step_breakdown_ms'], d['roofline']['frac']))
" || tail -5 gpurun_out/quick_$TAG.err
}
for w in $WHAT; do
case $w in
tests)
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_$TAG.log;;
testsall)
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_$TAG.log;;
smoke)
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_$TAG.log;;
bench)
  timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err;;
benchq)
  quick loops17
  quick borders9 --kernel borders --win-size 9 --pearson 0.15
  quick small7 --kernel loops_small --pearson 0.5;;
count)
  for cfg in "" "--kernel borders --win-size 9 --pearson 0.15" "--kernel loops_small --pearson 0.5"; do
    echo "count [$cfg]:"
    CHROMOSIGHT_B200_LIB=$PWD/chromosight_b200/libchromosight_b200_ablate.so CS_DEBUG_COUNT=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e $cfg 2>&1 >/dev/null | grep "pearson stats" | tail -1
  done;;
ablate)
  for m in ${MASKS:-0 8 32 40}; do
    echo -n "skip=$m: "
    CHROMOSIGHT_B200_LIB=$PWD/chromosight_b200/libchromosight_b200_ablate.so CS_DEBUG_SKIP=$m timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pearson ms %.3f' % d['step_breakdown_ms']['pearson'])"
  done;;
abk1)
  # A/B of the Pearson kernel variants: persistent vs one CTA per tile, 3 CTAs/SM build for small kernels
  echo "== default (persistent)"; quick loops17; quick borders9 --kernel borders --win-size 9 --pearson 0.15; quick small7 --kernel loops_small --pearson 0.5
  echo "== CS_PERSIST=0"; export CS_PERSIST=0; quick loops17; quick borders9 --kernel borders --win-size 9 --pearson 0.15; quick small7 --kernel loops_small --pearson 0.5; unset CS_PERSIST
  if [ -f chromosight_b200/libchromosight_b200_occ3.so ]; then
    echo "== occ3 build"; export CHROMOSIGHT_B200_LIB=$PWD/chromosight_b200/libchromosight_b200_occ3.so
    quick borders9 --kernel borders --win-size 9 --pearson 0.15; quick small7 --kernel loops_small --pearson 0.5; unset CHROMOSIGHT_B200_LIB
  fi;;
tilerows)
  # tile height of the Pearson kernel (ablate build reads CS_TILE_ROWS)
  export CHROMOSIGHT_B200_LIB=$PWD/chromosight_b200/libchromosight_b200_ablate.so
  for tr in 32 48 64 96; do
    export CS_TILE_ROWS=$tr; echo "== CS_TILE_ROWS=$tr"
    quick loops17; quick borders9 --kernel borders --win-size 9 --pearson 0.15; quick small7 --kernel loops_small --pearson 0.5
  done; unset CS_TILE_ROWS CHROMOSIGHT_B200_LIB;;
slab)
  timeout 300 python scripts/dbg_slab.py 4 1 2>&1 | tail -4
  CHROMOSIGHT_B200_LIB=$PWD/chromosight_b200/libchromosight_b200_ablate.so CS_DEBUG_COUNT=1 timeout 300 python scripts/dbg_slab.py 4 1 2>&1 | grep -i "pearson stats" | sort | uniq -c | tail -3;;
ncu7)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:pearson -s 3 -c 1 -o gpurun_out/prof7_$TAG -f \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --kernel loops_small --pearson 0.5 > gpurun_out/ncu_full7_$TAG.log 2>&1; echo "ncu full 7x7 rc=$?"; tail -2 gpurun_out/ncu_full7_$TAG.log;;
benchref)
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/benchref_$TAG.json 2> gpurun_out/benchref_$TAG.err; echo "benchref rc=$?"; cat gpurun_out/benchref_$TAG.json; nproc;;
ncu)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:pearson -s 3 -c 2 -o gpurun_out/prof_$TAG -f \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full_$TAG.log;;
esac
done
