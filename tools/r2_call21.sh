#!/bin/bash
# round 2, call 21: vector positional-encoding kernels, fully-masked chunk skip in the fused attention forward
set -u
OUT=gpurun_out/r2_call21
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run pytest_gpu 900 python -m pytest tests -m gpu -q -rs
run bench_tts 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
run bench_attn 400 python tools/bench_attn.py --out $OUT/bench_attn.json
grep -v "^$" $OUT/pytest_gpu.log | tail -8 | cut -c1-250
for f in bench_tts; do grep '"metric"' $OUT/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$f', d['value'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), 'roof', d.get('roofline', {}).get('frac'), d.get('gpu_launches_per_step'))
"; tail -2 $OUT/$f.log | cut -c1-200; done
tail -6 $OUT/bench_attn.log | cut -c1-220
