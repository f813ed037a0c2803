# round 2, GPU call 2: first run of the row-march binning kernel
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_bench_configs.py -m gpu -x -q 2>&1 | tail -15
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8
  python scripts/time_stages.py 32 32
  XCB200_NO_BIN_ROWS=1 python scripts/time_stages.py 32 32
  XC_NOISE=0 python scripts/time_stages.py 32 32
  XC_QUANT=8 python scripts/time_stages.py 32 32 ) 2>&1 | grep -v Warning | tee gpurun_out/r2_call2.txt
