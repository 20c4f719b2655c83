#!/bin/bash
# round 2, call 3: suite after gate / fp32 stream / trainer rewrite; full-depth error; default bench; other workloads
set -u
OUT=gpurun_out/r2_call3
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run pytest_gpu 900 python -m pytest tests -m gpu -q -rs -x
run err_layers 300 python tools/bf16_error_layers.py
run bench_tts 600 python bench.py --steps 20 --warmup 5
ST5_FFN_GATE=0 run bench_tts_nogate 300 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
ST5_FP32_STREAM=0 run bench_tts_nostream 300 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
ST5_GEMM_PAIR=0 run bench_tts_nopair 300 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
run bench_asr 600 python bench.py --workload asr --steps 10 --warmup 3
run bench_hifigan 400 python bench.py --workload hifigan --steps 20 --warmup 4
run bench_ragged 600 python bench.py --workload tts_ragged --steps 20
tail -3 $OUT/pytest_gpu.log; tail -25 $OUT/err_layers.log
for f in bench_tts bench_tts_nogate bench_tts_nostream bench_tts_nopair bench_asr bench_hifigan bench_ragged; do echo "== $f"; grep '"metric"' $OUT/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['value'], d['unit'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), 'roof', d.get('roofline', {}).get('frac'), d.get('modes'), d.get('roofline_conv0', {}).get('frac'), {k: v for k, v in d['config'].items() if k in ('cache_hit_rate','graphs_captured','hit_rate_second_pass','stream_utt_per_s_incl_captures')})
"; tail -2 $OUT/$f.log | head -1 | cut -c1-300; done
