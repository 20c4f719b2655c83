#!/bin/bash
# round 2, call 13: fused LN backward (smem strips + prefetch + dx column sums), dead-warp skip in the fused attention
# forward, ATen call-site attribution
set -u
OUT=gpurun_out/r2_call13
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run pytest_gpu 900 python -m pytest tests -m gpu -q -x -rs
run bench_ln 300 python tools/bench_ln.py --out $OUT/bench_ln.json
run bench_tts 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
ST5_FOLD_BIAS_GRAD=0 run bench_tts_nofold 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
run bench_attn 400 python tools/bench_attn.py --out $OUT/bench_attn.json
run glue_sites 400 python tools/glue_sites.py
cp gpurun_out/glue_sites_tts.txt $OUT/ 2>/dev/null
tail -6 $OUT/pytest_gpu.log; cat $OUT/bench_ln.log; tail -12 $OUT/bench_attn.log
for f in bench_tts bench_tts_nofold; do grep '"metric"' $OUT/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$f', d['value'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), 'roof', d.get('roofline', {}).get('frac'), d.get('gpu_launches_per_step'))
"; tail -2 $OUT/$f.log | cut -c1-200; done
head -70 $OUT/glue_sites.log | cut -c1-250
