#!/bin/bash
# Multi-GPU session (gpurun --gpus 8): configs 4 / 5 at their stated GPU counts and the weak /
# strong scaling lines of bench.py.  usage: bash scripts/gpu_multi.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { # label, nproc, args...
  local label=$1 np=$2; shift 2
  timeout 600 $TR --nproc-per-node $np --master-port $((29500 + RANDOM % 200)) bench.py --gpus $np "$@" \
      > gpurun_out/${label}_$TAG.json 2> gpurun_out/${label}_$TAG.err
  echo "$label rc=$?"; grep '^{' gpurun_out/${label}_$TAG.json | tail -1 | cut -c1-1500; tail -2 gpurun_out/${label}_$TAG.err | cut -c1-300
}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
run cfg4_n4 4 --config 4
run cfg5_n8 8 --config 5
run weak_n8 8 --steps 20 --no-cpu-baseline
run weak_n2 2 --steps 20 --no-cpu-baseline
run strong_n8 8 --steps 20 --no-cpu-baseline --no-e2e --scaling strong
