#!/bin/bash
# round 2, call 27: layer_norm extractor kernels + large-style reference vectors + joint pre-training update, then config 4
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_frontend_gpu.py tests/test_ref_pin_gpu.py -q -m gpu \
  -k "layer0_layer_norm or layer_norm_extractor or large_style or joint_pretraining" 2>&1 | tail -40 > gpurun_out/r2_ln_tests.txt
tail -5 gpurun_out/r2_ln_tests.txt
timeout 300 python bench.py --workload pretrain --steps 6 --warmup 3 > gpurun_out/r2_pretrain.json 2> gpurun_out/r2_pretrain.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r2_pretrain.json; tail -5 gpurun_out/r2_pretrain.err
