#!/bin/bash
# A/B of the k_lwa_cols scatter changes (packed LUT + one probe + LUT prefetch; half-warp word swap) + parity
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for rep in 1 2; do
for v in base noswap default; do
  if [ $v = default ]; then unset XCB200_LIB; else export XCB200_LIB=$PWD/xcontour_b200/libxcb200_$v.so; fi
  timeout 120 python scripts/time_stages.py 32 32
done; done
unset XCB200_LIB
XC_NOISE=0 timeout 120 python scripts/time_stages.py 32 32
XCB200_LIB=$PWD/xcontour_b200/libxcb200_base.so XC_NOISE=0 timeout 120 python scripts/time_stages.py 32 32
} > gpurun_out/r2_call16_ab.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_call16_tests.txt 2>&1
tail -3 gpurun_out/r2_call16_tests.txt
cat gpurun_out/r2_call16_ab.txt
