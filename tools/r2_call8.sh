#!/bin/bash
# round 2, call 8: full suite with the fused LayerNorm backward; per-shape GEMM table against the library; attention
# backward source-level profile after the saved-exponential rewrite; TTS bench
set -u
OUT=gpurun_out/r2_call8
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run pytest_gpu 900 python -m pytest tests -m gpu -q -rs
run probe_cublas 600 python tools/probe_cublas.py
run bench_tts 600 python bench.py --steps 20 --warmup 5 --no-parity --no-cpu-baseline
run attn_ncu 900 ncu --set full --import-source on --clock-control none -k regex:attn_fused --launch-skip 4 -c 4 -o $OUT/attn python tools/profile_attn.py
for i in 1 3; do
  ncu -i $OUT/attn.ncu-rep --page source --csv --print-source cuda,sass --launch-skip $i --launch-count 1 > $OUT/a_src.csv 2>/dev/null
  python tools/ncu_lines.py $OUT/a_src.csv 45 > $OUT/attn_lines_$i.txt 2>&1
  ncu -i $OUT/attn.ncu-rep --page source --csv --launch-skip $i --launch-count 1 > $OUT/a_sass.csv 2>/dev/null
  python tools/ncu_hot.py $OUT/a_sass.csv 70 > $OUT/attn_hot_$i.txt 2>&1
done
ncu -i $OUT/attn.ncu-rep --page raw --csv > $OUT/attn_raw.csv 2>/dev/null
rm -f $OUT/a_src.csv $OUT/a_sass.csv $OUT/attn.ncu-rep
tail -5 $OUT/pytest_gpu.log; cat $OUT/probe_cublas.log | cut -c1-220
grep '"metric"' $OUT/bench_tts.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['value'], d['unit'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'), 'roof', d.get('roofline', {}).get('frac'), d.get('gpu_launches_per_step'))
"
for i in 1 3; do head -40 $OUT/attn_lines_$i.txt; head -3 $OUT/attn_hot_$i.txt; done
