#!/bin/bash
# round 2, call 17: first-heads-only probabilities, deterministic criterion sums, graph-captured greedy text decoding,
# persistent synthesis graphs (seed undo), LN backward back at two CTAs
set -u
OUT=gpurun_out/r2_call17
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run pytest_gpu 900 python -m pytest tests -m gpu -q -rs
run bench_tts 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
ST5_WGRAD_SIDE=1 run bench_tts_side 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
run bench_asr 600 python bench.py --workload asr --steps 10 --warmup 3 --no-cpu-baseline
run bench_attn 400 python tools/bench_attn.py --out $OUT/bench_attn.json
run bench_inference 900 python tools/bench_inference.py --steps 150
run launches 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv python bench.py --profile-step --no-parity --no-cpu-baseline
cp gpurun_out/gemm_shapes.json $OUT/ 2>/dev/null
python tools/ncu_summary.py $OUT/launches.csv $OUT/gemm_shapes.json > $OUT/summary.txt 2>&1
grep -v "^$" $OUT/pytest_gpu.log | tail -12 | cut -c1-250
for f in bench_tts bench_tts_side bench_asr; do grep '"metric"' $OUT/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$f', d['value'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), 'roof', d.get('roofline', {}).get('frac'), d.get('gpu_launches_per_step'))
"; tail -2 $OUT/$f.log | cut -c1-200; done
tail -6 $OUT/bench_attn.log | cut -c1-250
tail -3 $OUT/bench_inference.log | cut -c1-1500
head -45 $OUT/summary.txt | cut -c1-150
