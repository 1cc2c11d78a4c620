#!/bin/bash
# One GPU-box session: parity tests, bench, micro-benchmark, ncu launch list + full capture.
# Usage (from the build container):  gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag> [what...]'
TAG=${1:-r1}; shift
WHAT=${@:-tests bench ncu}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
for w in $WHAT; do
case $w in
tests)
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log;;
smoke)
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_$TAG.log;;
bench)
  timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err;;
benchref)
  timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/benchref_$TAG.json 2> gpurun_out/benchref_$TAG.err; echo "benchref rc=$?"; cat gpurun_out/benchref_$TAG.json; nproc;;
ubench)
  for b in scripts/ubench/*; do [ -x $b ] && [ ! -d $b ] && case $b in *.cu) ;; *) timeout 120 $b;; esac; done > gpurun_out/ubench_$TAG.log 2>&1; cat gpurun_out/ubench_$TAG.log;;
ncu)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:pearson -s 3 -c 2 -o gpurun_out/prof_$TAG -f \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full_$TAG.log;;
esac
done
