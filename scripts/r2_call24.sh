#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for rep in 1 2; do
for v in base default; do
  if [ $v = default ]; then unset XCB200_LIB; else export XCB200_LIB=$PWD/xcontour_b200/libxcb200_$v.so; fi
  timeout 120 python scripts/time_stages.py 32 32
done; done
unset XCB200_LIB
} > gpurun_out/r2_call24_ab.txt 2>&1
cut -c1-170 gpurun_out/r2_call24_ab.txt
timeout 300 python -m pytest tests -m gpu -q -x -k "lwa or c4 or fused or smoke" 2>&1 | tail -3
