#!/bin/bash
# round 2, call 24: suite after the CTC switch in the eager criterion, ASR line
set -u
OUT=gpurun_out/r2_call24
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run pytest_gpu 900 python -m pytest tests -m gpu -q -rs
run bench_asr 600 python bench.py --workload asr --steps 10 --warmup 3 --no-cpu-baseline
grep -v "^$" $OUT/pytest_gpu.log | tail -8 | cut -c1-250
grep '"metric"' $OUT/bench_asr.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('asr', d['value'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'))
"
