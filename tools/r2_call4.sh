#!/bin/bash
# round 2, call 4: where the step goes now (launch list + shares), and source-level stall attribution of the fused
# attention kernels (ncu --set full --import-source on), summarised on the box so only text comes back
set -u
OUT=gpurun_out/r2_call4
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
# 0. full suite (no -x), the bf16 gradient diagnostic, the ASR capture diagnostic, ASR without the graph
run pytest_gpu 900 python -m pytest tests -m gpu -q -rs
run diag_grad 300 python tools/diag_grad_bf16.py
run diag_asr 300 python tools/diag_asr_capture.py
run bench_asr_nograph 600 python bench.py --workload asr --steps 6 --warmup 3 --no-graph
# 1. launch list of the default bench step
run launches 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv python bench.py --profile-step --no-parity --no-cpu-baseline
cp gpurun_out/gemm_shapes.json $OUT/ 2>/dev/null
python tools/ncu_summary.py $OUT/launches.csv $OUT/gemm_shapes.json > $OUT/summary.txt 2>&1
# 2. attention kernels, full sections with source
run attn_ncu 900 ncu --set full --import-source on --clock-control none -k regex:attn_fused --launch-skip 4 -c 4 -o $OUT/attn python tools/profile_attn.py
for i in 0 1 2 3; do
  ncu -i $OUT/attn.ncu-rep --page source --csv --print-source cuda,sass --launch-skip $i --launch-count 1 > $OUT/attn_src_$i.csv 2>/dev/null
  python tools/ncu_lines.py $OUT/attn_src_$i.csv 45 > $OUT/attn_lines_$i.txt 2>&1
  ncu -i $OUT/attn.ncu-rep --page source --csv --launch-skip $i --launch-count 1 > $OUT/attn_sass_$i.csv 2>/dev/null
  python tools/ncu_hot.py $OUT/attn_sass_$i.csv 60 > $OUT/attn_hot_$i.txt 2>&1
done
ncu -i $OUT/attn.ncu-rep --page raw --csv > $OUT/attn_raw.csv 2>/dev/null
rm -f $OUT/attn_src_*.csv $OUT/attn_sass_*.csv
tail -15 $OUT/pytest_gpu.log; cat $OUT/diag_grad.log | cut -c1-400; cat $OUT/diag_asr.log | tail -12; tail -3 $OUT/bench_asr_nograph.log | cut -c1-600
head -40 $OUT/summary.txt
for i in 0 1 2 3; do head -12 $OUT/attn_lines_$i.txt; head -3 $OUT/attn_hot_$i.txt; done
