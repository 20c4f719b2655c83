#!/bin/bash
# round 2, call 11: suite, attention table, TTS bench (+ parity leg and CPU baseline), launch list of the new step
set -u
OUT=gpurun_out/r2_call11
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run pytest_gpu 900 python -m pytest tests -m gpu -q -rs
run bench_attn 400 python tools/bench_attn.py --asr --out $OUT/bench_attn.json
run bench_tts 600 python bench.py --steps 20 --warmup 5
ST5_FOLD_RESGRAD=0 run bench_tts_nofold 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
run launches 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv python bench.py --profile-step --no-parity --no-cpu-baseline
cp gpurun_out/gemm_shapes.json $OUT/ 2>/dev/null
python tools/ncu_summary.py $OUT/launches.csv $OUT/gemm_shapes.json > $OUT/summary.txt 2>&1
tail -6 $OUT/pytest_gpu.log; cat $OUT/bench_attn.log
for f in bench_tts bench_tts_nofold; do grep '"metric"' $OUT/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$f', d['value'], d['unit'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), 'roof', d.get('roofline', {}).get('frac'), d['roofline'].get('gemm_ms_per_step'), d.get('gpu_launches_per_step'), d.get('modes'))
"; tail -2 $OUT/$f.log | cut -c1-200; done
head -60 $OUT/summary.txt
