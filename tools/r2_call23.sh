#!/bin/bash
# round 2, call 23: HiFi-GAN with the leaky-ReLU + padding + de-interleave of every convolution input in one launch
set -u
OUT=gpurun_out/r2_call23
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run pytest_gpu 900 python -m pytest tests -m gpu -q -rs
run bench_hifigan 400 python bench.py --workload hifigan --steps 5 --warmup 3 --no-cpu-baseline
run bench_tts 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
grep -v "^$" $OUT/pytest_gpu.log | tail -6 | cut -c1-250
grep -h '"metric"' $OUT/bench_hifigan.log | cut -c1-1400
grep '"metric"' $OUT/bench_tts.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('tts', d['value'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), 'roof', d.get('roofline', {}).get('frac'), d.get('gpu_launches_per_step'))
"
