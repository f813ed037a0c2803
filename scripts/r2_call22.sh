#!/bin/bash
# compute-sanitizer over the GPU suite: memcheck on everything that finishes in the time limit, racecheck on the
# tests that drive the round-2 kernels through unusual shapes and boundary rules
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( echo "## memcheck: pytest tests -m gpu (gather subprocess test excluded)"
  timeout 330 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "not nccl" 2>&1 | grep -vE "^=========\s*$" | tail -8
  echo "rc=${PIPESTATUS[0]}"
  echo "## racecheck: tests/test_gpu_bench_configs.py -k 'cartesian or row_march or lwa_f32 or streamer or c4_lwa_rows'"
  timeout 200 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests/test_gpu_bench_configs.py -q -x \
     -k "cartesian or row_march or lwa_f32 or streamer or c4_lwa_rows" 2>&1 | grep -vE "^=========\s*$" | tail -8
  echo "rc=${PIPESTATUS[0]}" ) > gpurun_out/r2_sanitizer_suite.txt 2>&1
cat gpurun_out/r2_sanitizer_suite.txt
