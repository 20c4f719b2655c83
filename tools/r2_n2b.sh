#!/bin/bash
# 2-GPU call: exchange check (eps = 1) and the N=2 bench under exchange / grouping variants
set -u
OUT=gpurun_out/r2_n2b
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
B="bench.py --gpus 2 --steps 20 --warmup 5 --no-parity --no-cpu-baseline"
run check_n2 400 $TR tools/check_n2.py
run n1 400 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
run shard_g1 400 $TR $B
ST5_STAGE_GROUP=3 run shard_g3 400 $TR $B
ST5_STAGE_GROUP=6 run shard_g6 400 $TR $B
ST5_STAGE_GROUP=3 run allreduce_g3 400 $TR $B --exchange allreduce
run allreduce_g1 400 $TR $B --exchange allreduce
ST5_OVERLAP_AR=0 run shard_nooverlap 400 $TR $B
grep -v "^\[W\|Warning\|warn" $OUT/check_n2.log | grep -v "^$" | tail -12
for f in n1 shard_g1 shard_g3 shard_g6 allreduce_g3 allreduce_g1 shard_nooverlap; do grep '"metric"' $OUT/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$f', d['value'], d['unit'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), d['config'].get('exchange'))
"; tail -1 $OUT/$f.log; done
