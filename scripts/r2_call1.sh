# round 2, GPU call 1: time the compile-time variants prepared at the end of round 1 + raw PCIe ceilings
mkdir -p gpurun_out
D=$PWD/xcontour_b200
( nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
  echo "== parity, variant all"
  XCB200_LIB=$D/libxcb200_all.so timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  for env in "" "XC_NOISE=0"; do
    echo "== field: ${env:-benchmark}"
    env $env python scripts/time_stages.py 32 32
    for v in own lut4k lean rowcnt own8 pair all; do env $env XCB200_LIB=$D/libxcb200_$v.so python scripts/time_stages.py 32 32; done
  done
  echo "== pcie ceiling, 1 GPU"
  python scripts/pcie_ceiling.py 512 ) 2>&1 | grep -v Warning | tee gpurun_out/r2_call1.txt
