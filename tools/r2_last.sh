#!/bin/bash
# round 2, last call: the suite, smoke and the default bench line at HEAD
set -u
OUT=gpurun_out/r2_last
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run pytest_gpu 900 python -m pytest tests -m gpu -q -rs
run smoke 300 python __graft_entry__.py smoke
run bench_tts 900 python bench.py --steps 20 --warmup 5
grep -v "^$" $OUT/pytest_gpu.log | tail -3 | cut -c1-200; tail -3 $OUT/smoke.log
grep '"metric"' $OUT/bench_tts.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('tts', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['gpu_launches_per_step'], d['modes']['bf16']['mel_rel_l2_vs_cpu_path'], d['modes']['parity']['mel_rel_l2_vs_cpu_path'])
"; tail -1 $OUT/bench_tts.log
