#!/bin/bash
# round 2, call 26: ragged-shape stream after pad_to_buckets pads the collater's `target` too (24 raw shapes -> 4 graphs)
set -u
OUT=gpurun_out/r2_call26
mkdir -p $OUT
( timeout 600 python bench.py --workload tts_ragged --steps 32 --warmup 3 --no-cpu-baseline ) > $OUT/bench_ragged.log 2>&1; echo "rc=$?" >> $OUT/bench_ragged.log
( timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -k "trainer" ) > $OUT/pytest_sub.log 2>&1
grep -h '"metric"' $OUT/bench_ragged.log | cut -c1-1500; tail -2 $OUT/pytest_sub.log
