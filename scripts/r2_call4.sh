# round 2, GPU call 4: parity of the fixed workspace + ncu --set full of the two new kernels
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_bench_configs.py -m gpu -x -q 2>&1 | tail -15
  ncu --set full --clock-control none --import-source on -k regex:"k_bin_rows|k_lwa_cols" -s 4 -c 2 -o gpurun_out/prof_r2_a python scripts/time_stages.py 32 32 > gpurun_out/p_r2_a.log 2>&1
  tail -1 gpurun_out/p_r2_a.log | cut -c1-200
  python scripts/ncu_summary.py gpurun_out/prof_r2_a.ncu-rep 30 > gpurun_out/r2_a_ncu_summary.txt 2>&1
  python scripts/ncu_source_lines.py k_bin_rows 45 gpurun_out/prof_r2_a.ncu-rep > gpurun_out/r2_a_src_bin_rows.txt 2>&1
  python scripts/ncu_source_lines.py k_lwa_cols 45 gpurun_out/prof_r2_a.ncu-rep > gpurun_out/r2_a_src_lwa_cols.txt 2>&1
  python scripts/time_stages.py 32 32 ) 2>&1 | grep -v Warning | tee gpurun_out/r2_call4.txt
