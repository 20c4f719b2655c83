#!/bin/bash
# round 2, call 20: LayerNorm backward with rows staged through a shared-memory ring (A/B), guided-head row constants,
# packed static inputs
set -u
OUT=gpurun_out/r2_call20
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run pytest_gpu 900 python -m pytest tests -m gpu -q -rs
run bench_ln 300 python tools/bench_ln.py --out $OUT/bench_ln.json
ST5_LN_BWD_RING=0 run bench_ln_regs 300 python tools/bench_ln.py --out $OUT/bench_ln_regs.json
run bench_tts 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
ST5_LN_BWD_RING=0 run bench_tts_regs 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
grep -v "^$" $OUT/pytest_gpu.log | tail -8 | cut -c1-250
grep ln_bwd $OUT/bench_ln.log; grep ln_bwd $OUT/bench_ln_regs.log
for f in bench_tts bench_tts_regs; do grep '"metric"' $OUT/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$f', d['value'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), 'roof', d.get('roofline', {}).get('frac'), d.get('gpu_launches_per_step'))
"; tail -2 $OUT/$f.log | cut -c1-200; done
