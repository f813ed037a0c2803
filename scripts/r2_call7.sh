# round 2, GPU call 7: LWA v4 quick parity + timing; C5 bench line (bounded)
mkdir -p gpurun_out
( timeout 120 python __graft_entry__.py --smoke-only 2>&1 | tail -2
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lwa or fused or workflow or lape or alternate" --durations=3 2>&1 | tail -8
  timeout 300 python -m pytest tests/test_gpu_bench_configs.py -m gpu -x -q -k "f32 or gather or gradient or cartesian or row_march or streamer or dlpack" --durations=3 2>&1 | tail -8
  python scripts/time_stages.py 32 32
  XC_NOISE=0 python scripts/time_stages.py 32 32
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_bin_rows|k_lwa_cols|k_minmax" -s 6 -c 3 python scripts/time_stages.py 32 32 2>&1 | grep -E "k_bin_rows|k_lwa_cols|k_minmax|duration|inst_executed|issue_active"
  timeout 240 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -3 ) 2>&1 | grep -v Warning | tee gpurun_out/r2_call7.txt
