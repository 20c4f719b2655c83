#!/bin/bash
# round 2, call 6: streaming attention (any length, clipped relative positions) first run; ASR under the graph
set -u
OUT=gpurun_out/r2_call6
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run pytest_flash 600 python -m pytest tests/test_a_ops_gpu.py -m gpu -q -k "streaming" -x
run pytest_gpu 900 python -m pytest tests -m gpu -q -rs
run bench_attn 400 python tools/bench_attn.py --asr --out $OUT/bench_attn.json
ST5_ATTN_FLASH=all run bench_attn_flash 400 python tools/bench_attn.py --out $OUT/bench_attn_flash.json
run diag_asr 300 python tools/diag_asr_capture.py
run bench_asr 600 python bench.py --workload asr --steps 10 --warmup 3
ST5_ATTN_FLASH=all run bench_tts_flash 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
tail -15 $OUT/pytest_flash.log; tail -5 $OUT/pytest_gpu.log; cat $OUT/bench_attn.log; cat $OUT/bench_attn_flash.log; tail -4 $OUT/diag_asr.log
for f in bench_asr bench_tts_flash; do grep '"metric"' $OUT/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['value'], d['unit'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), 'roof', d.get('roofline', {}).get('frac'), d.get('roofline_conv0'))
"; tail -3 $OUT/$f.log | cut -c1-300; done
