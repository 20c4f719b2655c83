#!/bin/bash
# round 2, call 12 (re-entry): full GPU suite, smoke, TTS bench (+parity leg, CPU baseline), ASR bench, attention table
set -u
OUT=gpurun_out/r2_call12
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run pytest_gpu 900 python -m pytest tests -m gpu -q -rs
run smoke 300 python __graft_entry__.py smoke
run bench_tts 600 python bench.py --steps 20 --warmup 5
run bench_asr 600 python bench.py --workload asr --steps 10 --warmup 3 --no-cpu-baseline
run bench_attn 400 python tools/bench_attn.py --asr --out $OUT/bench_attn.json
tail -6 $OUT/pytest_gpu.log; tail -4 $OUT/smoke.log; cat $OUT/bench_attn.log | tail -40
for f in bench_tts bench_asr; do grep '"metric"' $OUT/$f.log | cut -c1-1500; tail -2 $OUT/$f.log | cut -c1-300; done
