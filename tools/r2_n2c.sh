#!/bin/bash
# 2-GPU call: last validation after the side stream was scoped to the update (NCCL exchange check, N=2 line, trainer tests)
set -u
OUT=gpurun_out/r2_n2c
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523"
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run check_n2 400 $TR tools/check_n2.py
run bench_n2_shard 500 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-parity --no-cpu-baseline
run pytest_sub 600 python -m pytest tests/test_model_gpu.py tests/test_frontend_gpu.py -m gpu -q -k "trainer or synthesis or speech_to_text or greedy"
grep -v "^\[W\|Warning\|warn\|\*\*\*\|OMP_NUM\|^$" $OUT/check_n2.log | tail -3
grep '"metric"' $OUT/bench_n2_shard.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('n2', d['value'], d['unit'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), d['config'].get('exchange'))
"; tail -3 $OUT/pytest_sub.log | cut -c1-200
