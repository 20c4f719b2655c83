#!/bin/bash
# round 2, call 2: un-gated suite (0 skipped expected), full-depth error attribution, inference bench, glue attribution
set -u
OUT=gpurun_out/r2_call2
mkdir -p $OUT
run() { local name=$1 t=$2; shift 2; ( timeout $t "$@" ) > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run pytest_gpu 900 python -m pytest tests -m gpu -q -rs
run err_layers 300 python tools/bf16_error_layers.py
run err_layers_22 200 python tools/bf16_error_layers.py --layers 2 2
run bench_inference 600 python tools/bench_inference.py --steps 100
run glue 300 python tools/profile_glue.py
cp gpurun_out/*.txt gpurun_out/*.json $OUT/ 2>/dev/null
tail -3 $OUT/pytest_gpu.log; cat $OUT/err_layers.log | tail -30; tail -5 $OUT/bench_inference.log; tail -4 $OUT/glue.log
