#!/bin/bash
# round 2, call 14: LN backward fix, three-kernel CTC, pre-training criteria, mask-cast cache, direct table gradient
set -u
OUT=gpurun_out/r2_call14
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run pytest_gpu 900 python -m pytest tests -m gpu -q -rs
run bench_tts 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
ST5_FOLD_BIAS_GRAD=0 run bench_tts_nofold 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
run bench_asr 600 python bench.py --workload asr --steps 10 --warmup 3 --no-cpu-baseline
run launches_asr 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_asr.csv python bench.py --workload asr --profile-step --no-cpu-baseline
python tools/ncu_summary.py $OUT/launches_asr.csv > $OUT/summary_asr.txt 2>&1
tail -8 $OUT/pytest_gpu.log
for f in bench_tts bench_tts_nofold bench_asr; do grep '"metric"' $OUT/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$f', d['value'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), 'roof', d.get('roofline', {}).get('frac'), d.get('gpu_launches_per_step'))
"; tail -2 $OUT/$f.log | cut -c1-200; done
head -40 $OUT/summary_asr.txt | cut -c1-150
