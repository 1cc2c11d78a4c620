#!/bin/bash
# A/B of the host staging knobs inside ONE box (boxes differ): wall ms of the e2e call
for rep in 1 2; do
for cfg in "spawn 6 32" "pool 8 16" "pool 4 16" "pool 2 16" "spawn 4 32" "spawn 3 32" "spawn 2 32" "pool 4 32" "pool 6 8" "spawn 1 32"; do
  set -- $cfg
  echo -n "mode $1 threads $2 chunk $3: "
  CS_COPY_MODE=$1 CS_COPY_THREADS=$2 CS_STAGE_CHUNK_MB=$3 SLABS=${SLABS:-12,8} TRACE_SLABS=12 timeout 100 python scripts/pipeline_sweep.py 2>/dev/null | tr '\n' ' '
  echo
done; done
