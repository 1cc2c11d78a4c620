#!/bin/bash
# strong scaling of ONE 200k-bin chromosome in row slabs (bench.py --scaling strong) at N = 1, 2, 4 [8]
TAG=${1:-r2}; shift
mkdir -p gpurun_out
for np in ${@:-1 2 4}; do
  if [ $np = 1 ]; then
    timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-e2e > gpurun_out/strong_n1_$TAG.json 2> gpurun_out/strong_n1_$TAG.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $np --master-port $((29500 + RANDOM % 200)) \
      bench.py --gpus $np --steps 20 --no-cpu-baseline --no-e2e --scaling strong > gpurun_out/strong_n${np}_$TAG.json 2> gpurun_out/strong_n${np}_$TAG.err
  fi
  grep '^{' gpurun_out/strong_n${np}_$TAG.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=%d value %.3e ms/step %.3f per-rank %s breakdown %s cand %s gathered %s' % (d['n_gpus'], d['value'], d['ms_per_step'], [round(x,3) for x in d.get('per_rank_ms_per_step',[])], d['step_breakdown_ms'], d['candidates_per_map'], d['candidates_gathered']))"
done
