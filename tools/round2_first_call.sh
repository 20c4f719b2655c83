#!/bin/bash
# First GPU call of round 2 (one B200): everything that was written at the end of round 1 without GPU time gets its
# first run, and the CTA-pair GEMM gets measured. Usage (from this container):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# All output lands in gpurun_out/r2_first/ (merged back by gpurun); the last lines printed are a one-screen summary.
set -u
OUT=gpurun_out/r2_first
mkdir -p $OUT
run() {  # name, timeout, command...
  local name=$1 t=$2; shift 2
  ( timeout $t "$@" ) > $OUT/$name.log 2>&1
  echo "rc=$?" >> $OUT/$name.log
}
# 1. the validated suite is still green (regression guard: GEMM template change, model forward dispatch change)
run pytest_gpu 600 python -m pytest tests -m gpu -x -q
# 2. first runs of the gated tests, each on its own so one failure does not hide the others
ST5_TEST_UNFUSED=1 run gated_unfused 120 python -m pytest tests/test_a_ops_gpu.py -k tensor_core_attention -q
ST5_TEST_T2T=1 run gated_t2t 120 python -m pytest tests/test_model_gpu.py -k text_to_text -q
ST5_TEST_CONV0=1 run gated_conv0 120 python -m pytest tests/test_a_ops_gpu.py -k conv0 -q
ST5_TEST_FRONTEND=1 run gated_frontend 900 python -m pytest tests/test_frontend_gpu.py -q   # no -x: every draft gets its verdict
# 3. CTA-pair GEMM: bit-exactness against the single-CTA kernel + per-shape timings of both (what the ST5_TEST_PAIR
#    test wraps)
run pair_check 300 python tools/check_gemm_pair.py
# 4. the step with and without the pair GEMM (same box, back to back)
run bench_single 400 python bench.py --steps 20 --warmup 5
ST5_GEMM_PAIR=1 run bench_pair 400 python bench.py --steps 20 --warmup 5
ST5_GEMM_PAIR=2 run bench_pair5 400 python bench.py --steps 20 --warmup 5   # same kernel, 5-stage operand ring
# 6. launch list of the faster of the two is taken in a follow-up call (ncu replays are slow); summary:
echo "==== summary"
for f in pytest_gpu gated_unfused gated_t2t gated_conv0 gated_frontend; do
  echo "$f: $(grep -E 'passed|failed|error' $OUT/$f.log | tail -1) $(tail -1 $OUT/$f.log)"
done
grep -E "PAIR GEMM|shape [0-9]+: single" $OUT/pair_check.log | tail -14
head -12 gpurun_out/glue_profile.txt 2>/dev/null
for f in bench_single bench_pair bench_pair5; do
  python - "$OUT/$f.log" <<'PY'
import json, sys
for line in open(sys.argv[1]):
    line = line.strip()
    if line.startswith("{") and '"metric"' in line:
        d = json.loads(line)
        print(sys.argv[1].split("/")[-1], d["value"], d["unit"], d["ms_per_step"], "ms/step e2e", d.get("e2e", {}).get("value"),
              "roofline", d.get("roofline", {}).get("frac"))
PY
done
