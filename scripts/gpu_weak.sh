#!/bin/bash
# weak scaling line of bench.py at the given rank counts (device-timed value only)
TAG=${1:-r2}; shift
mkdir -p gpurun_out
for np in ${@:-1 2}; do
  if [ $np = 1 ]; then
    timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-e2e > gpurun_out/weak_n1_$TAG.json 2> gpurun_out/weak_n1_$TAG.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $np --master-port $((29500 + RANDOM % 200)) \
      bench.py --gpus $np --steps 30 --no-cpu-baseline --no-e2e > gpurun_out/weak_n${np}_$TAG.json 2> gpurun_out/weak_n${np}_$TAG.err
  fi
  echo "N=$np rc=$?"; tail -2 gpurun_out/weak_n${np}_$TAG.err | cut -c1-300
  grep '^{' gpurun_out/weak_n${np}_$TAG.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=%d value %.3e ms/step %.3f breakdown %s launches %s per-rank %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['step_breakdown_ms'], d['gpu_launches'], [round(x,3) for x in d['per_rank_ms_per_step']]))"
done
