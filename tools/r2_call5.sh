#!/bin/bash
# round 2, call 5: saved-exponential attention (sign-bit dropout, fp32 delta), ASR capture on one stream
set -u
OUT=gpurun_out/r2_call5
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run pytest_gpu 900 python -m pytest tests -m gpu -q -rs
run diag_grad 300 python tools/diag_grad_bf16.py
run diag_asr 300 python tools/diag_asr_capture.py
run bench_attn 300 python tools/bench_attn.py --asr --out $OUT/bench_attn.json
run bench_tts 600 python bench.py --steps 20 --warmup 5
run bench_asr 600 python bench.py --workload asr --steps 10 --warmup 3
tail -12 $OUT/pytest_gpu.log; cut -c1-330 $OUT/diag_grad.log; tail -8 $OUT/diag_asr.log; cat $OUT/bench_attn.log
for f in bench_tts bench_asr; do grep '"metric"' $OUT/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['value'], d['unit'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), 'roof', d.get('roofline', {}).get('frac'), d.get('modes'), d.get('roofline_conv0'))
"; tail -3 $OUT/$f.log | cut -c1-300; done
