#!/bin/bash
# round 2, call 9: attention backward with TMA-staged exponentials, GEMM epilogue prefetch, weight-gradient split-K with
# the L2 reduce
set -u
OUT=gpurun_out/r2_call9
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run pytest_gpu 900 python -m pytest tests -m gpu -q -rs
run bench_attn 400 python tools/bench_attn.py --asr --out $OUT/bench_attn.json
run bench_tts 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
ST5_WGRAD_SPLITK=0 run bench_tts_nosplitk 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
run probe_cublas 600 python tools/probe_cublas.py
tail -8 $OUT/pytest_gpu.log; cat $OUT/bench_attn.log
for f in bench_tts bench_tts_nosplitk; do grep '"metric"' $OUT/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$f', d['value'], d['unit'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), 'roof', d.get('roofline', {}).get('frac'), d['roofline'].get('gemm_ms_per_step'), d.get('gpu_launches_per_step'))
"; tail -2 $OUT/$f.log | cut -c1-300; done
tail -8 $OUT/probe_cublas.log | cut -c1-200
