#!/bin/bash
# Weak scaling of the c3 training step (512 rows per GPU): tools/weak_scaling.sh "1 2 4 8" -> gpurun_out/weak_c3_<N>gpu.json
for n in $1; do
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline --no-reference-iteration \
      > gpurun_out/weak_c3_${n}gpu.json 2> gpurun_out/weak_c3_${n}gpu.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 \
      bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/weak_c3_${n}gpu.json 2> gpurun_out/weak_c3_${n}gpu.err
  fi
  python -c "import json; d=json.loads(open('gpurun_out/weak_c3_${n}gpu.json').read()); print(d['n_gpus'], d['scaling'], round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['ms_per_step'],3))"
done
