#!/bin/bash
# N-GPU bench lines (as the driver launches them); usage: bench_tools/run_scaling.sh 8 2
for N in "$@"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N)) \
      bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
  echo "N=$N rc=$?"
  python -c "
import json
d = json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
done
