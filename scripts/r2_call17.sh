#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for rep in 1 2; do
for v in base v1 v2 v3 v4; do
  export XCB200_LIB=$PWD/xcontour_b200/libxcb200_$v.so
  timeout 120 python scripts/time_stages.py 32 32
done; done
for v in base v1 v3; do
  export XCB200_LIB=$PWD/xcontour_b200/libxcb200_$v.so
  XC_NOISE=0 timeout 120 python scripts/time_stages.py 32 32
done
} > gpurun_out/r2_call17_ab.txt 2>&1
cat gpurun_out/r2_call17_ab.txt
