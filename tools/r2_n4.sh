#!/bin/bash
# 4-GPU call: the bench line at N=4 (default exchange = shard)
set -u
OUT=gpurun_out/r2_n4
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29525"
( timeout 600 $TR bench.py --gpus 4 --steps 20 --warmup 5 --no-parity --no-cpu-baseline ) > $OUT/bench_n4.log 2>&1; echo "rc=$?" >> $OUT/bench_n4.log
grep '"metric"' $OUT/bench_n4.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('n4', d['value'], d['unit'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), d['config'].get('exchange'))
"; tail -2 $OUT/bench_n4.log | cut -c1-200
