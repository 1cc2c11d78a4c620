#!/bin/bash
# second multi-GPU session of round 2: config 5 at 8 ranks, weak scaling value at 8 ranks
TAG=${1:-r2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { local label=$1 np=$2; shift 2
  timeout 500 $TR --nproc-per-node $np --master-port $((29500 + RANDOM % 200)) bench.py --gpus $np "$@" \
      > gpurun_out/${label}_$TAG.json 2> gpurun_out/${label}_$TAG.err
  echo "$label rc=$?"; grep '^{' gpurun_out/${label}_$TAG.json | tail -1 | cut -c1-900; tail -1 gpurun_out/${label}_$TAG.err | cut -c1-200; }
run cfg5_n8 8 --config 5
run cfg4_n4 4 --config 4
run weak_n8 8 --steps 30 --no-cpu-baseline --no-e2e
