# round 2, GPU call 6: LWA v3, TMA min/max A/B, C5 through the fused batch; quick parity; ncu source lines
mkdir -p gpurun_out
( timeout 120 python __graft_entry__.py --smoke-only 2>&1 | tail -2
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lwa or fused or workflow or levels or lape" --durations=5 2>&1 | tail -12
  python scripts/time_stages.py 32 32
  XCB200_NO_BULK=1 python scripts/time_stages.py 32 32
  XC_NOISE=0 python scripts/time_stages.py 32 32
  timeout 600 python -m pytest tests/test_gpu_bench_configs.py -m gpu -x -q -k "c5 or gather or cartesian" --durations=5 2>&1 | tail -12
  timeout 300 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -2
  ncu --set full --clock-control none --import-source on -k regex:"k_lwa_cols" -s 2 -c 1 -o gpurun_out/prof_r2_b python scripts/time_stages.py 32 32 > gpurun_out/p_r2_b.log 2>&1
  python scripts/ncu_summary.py gpurun_out/prof_r2_b.ncu-rep 25 > gpurun_out/r2_b_ncu_summary.txt 2>&1
  python scripts/ncu_source_lines.py k_lwa_cols 60 gpurun_out/prof_r2_b.ncu-rep > gpurun_out/r2_b_src_lwa_cols.txt 2>&1 ) 2>&1 | grep -v Warning | tee gpurun_out/r2_call6.txt
