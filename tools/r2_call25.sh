#!/bin/bash
# round 2, call 25: two A/B runs with existing switches -- streaming attention forward for every shape (2 CTAs / SM without
# relative positions) and bias-gradient column sums kept out of the LayerNorm backward (they ride on the side stream now)
set -u
OUT=gpurun_out/r2_call25
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
ST5_ATTN_FLASH=all run bench_attn_flash 400 python tools/bench_attn.py --out $OUT/bench_attn_flash.json
ST5_ATTN_FLASH=all run bench_tts_flash 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
ST5_FOLD_BIAS_GRAD=0 run bench_tts_nofold 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
run bench_tts 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
tail -6 $OUT/bench_attn_flash.log | cut -c1-220
for f in bench_tts bench_tts_flash bench_tts_nofold; do grep '"metric"' $OUT/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$f', d['value'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), d.get('gpu_launches_per_step'))
"; tail -1 $OUT/$f.log; done
