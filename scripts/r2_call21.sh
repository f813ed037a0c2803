#!/bin/bash
# compute-sanitizer memcheck over the parity tests that drive the round-2 kernels through unusual shapes and
# boundary rules (Cartesian / X-Z planes, ghost-cell rules, small planes, fp32 LWA output, d/dA coordinates)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_bench_configs.py -q -x \
  -k "cartesian or row_march or lwa_f32 or gradient or interp_to_coords or streamer" > gpurun_out/r2_sanitizer_tests.txt 2>&1
echo "rc=$?" >> gpurun_out/r2_sanitizer_tests.txt
grep -E "ERROR SUMMARY|passed|failed|rc=" gpurun_out/r2_sanitizer_tests.txt | tail -5
