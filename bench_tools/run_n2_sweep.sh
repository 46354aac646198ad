#!/bin/bash
# all-reduce bucketing / NCCL channel sweep at N=2
run() {
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'])"
}
AIMB200_BUCKET_BLOCKS=3 run "bucket3"
AIMB200_BUCKET_BLOCKS=12 run "bucket12"
AIMB200_BUCKET_BLOCKS=3 NCCL_MAX_CTAS=4 run "bucket3_ctas4"
AIMB200_BUCKET_BLOCKS=12 NCCL_MAX_CTAS=4 run "bucket12_ctas4"
AIMB200_BUCKET_BLOCKS=1 NCCL_MAX_CTAS=2 run "bucket1_ctas2"
