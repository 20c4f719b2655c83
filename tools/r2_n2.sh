#!/bin/bash
# 2-GPU call: exchange correctness at N=2 on NCCL and the N=2 bench line for both exchange modes
set -u
OUT=gpurun_out/r2_n2
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run check_n2 400 $TR tools/check_n2.py
run bench_n2_shard 500 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-parity --no-cpu-baseline
run bench_n2_allreduce 500 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-parity --no-cpu-baseline --exchange allreduce
run bench_n1 400 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
grep -v "^\[W\|Warning\|warn" $OUT/check_n2.log | tail -14
for f in bench_n1 bench_n2_shard bench_n2_allreduce; do grep '"metric"' $OUT/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$f', d['value'], d['unit'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), d['config'].get('exchange'))
"; tail -2 $OUT/$f.log | cut -c1-200; done
