#!/bin/bash
# round 2, call 7: streaming attention after the window fix; graph-timed attention table; source-level GEMM profile;
# launch list of the ASR step
set -u
OUT=gpurun_out/r2_call7
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run pytest_flash 600 python -m pytest tests/test_a_ops_gpu.py -m gpu -q -k "streaming"
run bench_attn 400 python tools/bench_attn.py --asr --out $OUT/bench_attn.json
ST5_ATTN_FLASH=all run bench_attn_flash 400 python tools/bench_attn.py --out $OUT/bench_attn_flash.json
run gemm_ncu 900 ncu --set full --import-source on --clock-control none -k regex:gemm_bf16 -o $OUT/gemm python tools/profile_gemm.py
ncu -i $OUT/gemm.ncu-rep --page raw --csv > $OUT/gemm_raw.csv 2>/dev/null
for i in 1 3 5 7 9 11 13 15 17; do
  ncu -i $OUT/gemm.ncu-rep --page source --csv --print-source cuda,sass --launch-skip $i --launch-count 1 > $OUT/g_src.csv 2>/dev/null
  python tools/ncu_lines.py $OUT/g_src.csv 40 > $OUT/gemm_lines_$i.txt 2>&1
  ncu -i $OUT/gemm.ncu-rep --page source --csv --launch-skip $i --launch-count 1 > $OUT/g_sass.csv 2>/dev/null
  python tools/ncu_hot.py $OUT/g_sass.csv 50 > $OUT/gemm_hot_$i.txt 2>&1
done
rm -f $OUT/g_src.csv $OUT/g_sass.csv $OUT/gemm.ncu-rep
run asr_launches 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/asr_launches.csv python bench.py --workload asr --profile-step --no-cpu-baseline
python tools/ncu_summary.py $OUT/asr_launches.csv > $OUT/asr_summary.txt 2>&1
tail -6 $OUT/pytest_flash.log; cat $OUT/bench_attn.log; cat $OUT/bench_attn_flash.log; head -45 $OUT/asr_summary.txt
for i in 1 3 5 7; do head -14 $OUT/gemm_lines_$i.txt; head -3 $OUT/gemm_hot_$i.txt; done
