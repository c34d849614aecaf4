#!/bin/bash
# compute-sanitizer over the product kernels.  usage (on the GPU box): scripts/sanitize.sh <out dir> [tests]
#   1. smoke() (flow kernels, fused populate turn, accept, three training epochs) under memcheck,
#      synccheck, initcheck and racecheck;
#   2. with "tests": the GPU parity tests of every kernel family under memcheck, and the shared-memory
#      heavy ones (training, populate / accept, non-affine tail) under racecheck (38 minutes on a B200).
out=${1:-gpurun_out/sanitizer}
mkdir -p "$out"
for tool in memcheck synccheck initcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 \
      python -c "import __graft_entry__ as e; e.smoke()" > "$out/smoke_$tool.log" 2>&1
  echo "smoke $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok" "$out/smoke_$tool.log" | tail -3
done
[ "$2" = tests ] || exit 0
T="tests/test_gpu_flow.py tests/test_gpu_populate.py tests/test_gpu_train.py tests/test_gpu_zz_tail_accumulate.py tests/test_gpu_importance.py tests/test_gpu_c3_nsf.py"
timeout 2400 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 \
    python -m pytest $T -x -q -k "not large_batch" > "$out/tests_memcheck.log" 2>&1
echo "tests memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" "$out/tests_memcheck.log" | tail -3
timeout 2400 compute-sanitizer --tool racecheck --print-limit 20 --error-exitcode 9 \
    python -m pytest tests/test_gpu_train.py tests/test_gpu_populate.py tests/test_gpu_zz_tail_accumulate.py -x -q > "$out/tests_racecheck.log" 2>&1
echo "tests racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" "$out/tests_racecheck.log" | tail -3
