#!/bin/bash
# round 2, final measurement call: suite, smoke, the full bench line (both precision modes, CPU baseline), the reference arm,
# the ASR / HiFi-GAN / ragged lines, launch list with DRAM traffic of the final step, full ncu captures of the new kernels
set -u
OUT=gpurun_out/r2_final
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run pytest_gpu 900 python -m pytest tests -m gpu -q -rs
run smoke 300 python __graft_entry__.py smoke
run bench_tts 900 python bench.py --steps 20 --warmup 5
run bench_ref 600 python bench.py --impl reference --steps 3 --warmup 1
run bench_asr 600 python bench.py --workload asr --steps 10 --warmup 3
run bench_ragged 600 python bench.py --workload tts_ragged --steps 32 --warmup 3 --no-cpu-baseline
run bench_hifigan 400 python bench.py --workload hifigan --steps 5 --warmup 3 --no-cpu-baseline
run launches 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv python bench.py --profile-step --no-parity --no-cpu-baseline
cp gpurun_out/gemm_shapes.json $OUT/ 2>/dev/null
python tools/ncu_summary.py $OUT/launches.csv $OUT/gemm_shapes.json > $OUT/summary.txt 2>&1
python tools/ncu_traffic.py $OUT/launches.csv $OUT/gemm_traffic.json > $OUT/traffic.log 2>&1
run ln_ncu 600 ncu --set full --clock-control none -k regex:"ln_bwd_fused|ln_fwd|colsum_vec" --launch-skip 40 -c 6 -o $OUT/ln python tools/bench_ln.py
ncu -i $OUT/ln.ncu-rep --page raw --csv > $OUT/ln_raw.csv 2>/dev/null
python tools/ncu_raw_pick.py $OUT/ln_raw.csv > $OUT/ln_ncu_full.txt 2>&1
rm -f $OUT/ln.ncu-rep $OUT/ln_raw.csv
run bench_attn 400 python tools/bench_attn.py --asr --out $OUT/bench_attn.json
grep -v "^$" $OUT/pytest_gpu.log | tail -6 | cut -c1-250; tail -3 $OUT/smoke.log
for f in bench_tts bench_asr bench_ragged bench_hifigan bench_ref; do grep -h '"metric"\|"impl"' $OUT/$f.log | cut -c1-2600; tail -1 $OUT/$f.log; done
head -30 $OUT/summary.txt | cut -c1-150; cat $OUT/traffic.log | cut -c1-400; cat $OUT/ln_ncu_full.txt | cut -c1-400; tail -4 $OUT/bench_attn.log | cut -c1-200
