#!/bin/bash
# round 2, call 28: where does the memory of the config-4 update go (per module / per micro-batch), 2 updates
mkdir -p gpurun_out
ST5_MEMLOG=1 timeout 200 python bench.py --workload pretrain --steps 1 --warmup 1 > gpurun_out/r2_pretrain_mem.json 2> gpurun_out/r2_pretrain_mem.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2_pretrain_mem.json; grep -v "^  File\|^    " gpurun_out/r2_pretrain_mem.err | tail -60
