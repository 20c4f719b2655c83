#!/bin/bash
# 2-GPU call after the side-stream column sums: exchange correctness on NCCL + the N=2 bench line
set -u
OUT=gpurun_out/r2_n2b
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run check_n2 400 $TR tools/check_n2.py
run bench_n2_shard 500 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-parity --no-cpu-baseline
grep -v "^\[W\|Warning\|warn\|\*\*\*\|OMP_NUM" $OUT/check_n2.log | tail -12
grep '"metric"' $OUT/bench_n2_shard.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('n2', d['value'], d['unit'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), d['config'].get('exchange'))
"; tail -2 $OUT/bench_n2_shard.log | cut -c1-200
