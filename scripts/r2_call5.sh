# round 2, GPU call 5: branch-free k_lwa_cols + bin_rows vote fix: parity subset, timing, light ncu
mkdir -p gpurun_out
( timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python __graft_entry__.py --smoke-only 2>&1 | tail -4
  timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
  timeout 900 python -m pytest tests/test_gpu_bench_configs.py -m gpu -x -q -k "c4 or cartesian or row_march or gradient" 2>&1 | tail -8
  python scripts/time_stages.py 32 32
  XC_NOISE=0 python scripts/time_stages.py 32 32
  XC_QUANT=8 python scripts/time_stages.py 32 32
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_bin_rows|k_lwa_cols" -s 4 -c 2 python scripts/time_stages.py 32 32 2>&1 | grep -E "k_bin_rows|k_lwa_cols|duration|inst_executed|issue_active|bank_conflicts|wavefronts|warps_active" ) 2>&1 | grep -v Warning | tee gpurun_out/r2_call5.txt
