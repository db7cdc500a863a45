#!/bin/bash
# Strong scaling of the c3 training step at a fixed GLOBAL minibatch of 4096 rows:
#   tools/strong_scaling.sh "1 2 4"    -> gpurun_out/strong_c3_<N>gpu.json, one bench line per N
for n in $1; do
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps 30 --warmup 5 --global-rows 4096 --no-cpu-baseline --no-reference-iteration \
      > gpurun_out/strong_c3_${n}gpu.json 2> gpurun_out/strong_c3_${n}gpu.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 30 --warmup 5 --global-rows 4096 --no-cpu-baseline \
      > gpurun_out/strong_c3_${n}gpu.json 2> gpurun_out/strong_c3_${n}gpu.err
  fi
  python -c "import json,sys; d=json.loads(open('gpurun_out/strong_c3_${n}gpu.json').read()); print(d['n_gpus'], d['scaling'], round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['ms_per_step'],3))"
done
