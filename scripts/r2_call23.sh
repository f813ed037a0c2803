#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2_gpu_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_gpu_tests.txt
tail -15 gpurun_out/r2_gpu_tests.txt
timeout 100 python __graft_entry__.py --smoke-only 2>&1 | tail -2
