#!/bin/bash
# 8-GPU call: the bench line at N=8 (default exchange = shard) -- validation of the scaling path the driver runs
set -u
OUT=gpurun_out/r2_n8
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
( timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-parity --no-cpu-baseline ) > $OUT/bench_n8.log 2>&1; echo "rc=$?" >> $OUT/bench_n8.log
grep '"metric"' $OUT/bench_n8.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('n8', d['value'], d['unit'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), d['config'].get('exchange'))
"; tail -3 $OUT/bench_n8.log | cut -c1-300
