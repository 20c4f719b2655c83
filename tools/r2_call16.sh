#!/bin/bash
# round 2, call 16: LN backward at 3 CTAs/SM (A/B), persistent synthesis graphs with chunked stop reads, BN gradients in place
set -u
OUT=gpurun_out/r2_call16
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run pytest_gpu 900 python -m pytest tests -m gpu -q -rs
run bench_ln 300 python tools/bench_ln.py --out $OUT/bench_ln.json
ST5_LN_BWD_2CTA=1 run bench_ln_2cta 300 python tools/bench_ln.py --out $OUT/bench_ln_2cta.json
run bench_tts 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
ST5_LN_BWD_2CTA=1 run bench_tts_2cta 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
run bench_inference 600 python tools/bench_inference.py --steps 150
grep -v "^$" $OUT/pytest_gpu.log | tail -12 | cut -c1-250
grep ln_bwd $OUT/bench_ln.log; grep ln_bwd $OUT/bench_ln_2cta.log
for f in bench_tts bench_tts_2cta; do grep '"metric"' $OUT/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$f', d['value'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), 'roof', d.get('roofline', {}).get('frac'), d.get('gpu_launches_per_step'))
"; tail -2 $OUT/$f.log | cut -c1-200; done
tail -3 $OUT/bench_inference.log | cut -c1-700
