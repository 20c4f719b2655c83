#!/bin/bash
# round 2, call 22: bias-gradient column sums and the table gradient on the weight-gradient stream (A/B), ncu --set full
# of the attention kernels at the final state
set -u
OUT=gpurun_out/r2_call22
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run pytest_gpu 900 python -m pytest tests -m gpu -q -rs
run bench_tts 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
ST5_SIDE_SMALL=0 run bench_tts_main 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
run attn_ncu 900 ncu --set full --clock-control none -k regex:"attn_fused|attn_delta|attn_dqp" --launch-skip 8 -c 12 -o $OUT/attn python tools/profile_attn.py
ncu -i $OUT/attn.ncu-rep --page raw --csv > $OUT/attn_raw.csv 2>/dev/null
python tools/ncu_raw_pick.py $OUT/attn_raw.csv > $OUT/attn_ncu_full.txt 2>&1
rm -f $OUT/attn.ncu-rep $OUT/attn_raw.csv
grep -v "^$" $OUT/pytest_gpu.log | tail -6 | cut -c1-250
for f in bench_tts bench_tts_main; do grep '"metric"' $OUT/$f.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$f', d['value'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), 'roof', d.get('roofline', {}).get('frac'), d.get('gpu_launches_per_step'))
"; tail -2 $OUT/$f.log | cut -c1-200; done
cat $OUT/attn_ncu_full.txt | cut -c1-330
