# Round-end evidence run (1 GPU): full GPU suite, sanitizers, launch list, full ncu capture of the hot kernels, bench lines.
# Every step writes straight to its own file under gpurun_out/ and runs under its own timeout.
mkdir -p gpurun_out
R=${1:-r2}
timeout 900 python -u -m pytest tests -m gpu -q --durations=10 > gpurun_out/${R}_gpu_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/${R}_gpu_tests.txt
( echo "## racecheck, default kernels (k_minmax_bulk, k_bin_rows, k_lwa_cols, k_scan_epilogue)"
  timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis python __graft_entry__.py --smoke-only 2>&1 | grep -E "Race reported|RACECHECK SUMMARY|smoke ok" | sed -E 's/\+0x[0-9a-f]+//g' | sort | uniq -c
  echo "## racecheck, fallbacks (XCB200_NO_BIN_ROWS=1 XCB200_NO_LWA_COLS=1 XCB200_NO_BULK=1: k_hist, k_lwa_fx, k_minmax_partial)"
  XCB200_NO_BIN_ROWS=1 XCB200_NO_LWA_COLS=1 XCB200_NO_BULK=1 timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis python __graft_entry__.py --smoke-only 2>&1 | grep -E "Race reported|RACECHECK SUMMARY|smoke ok" | sed -E 's/\+0x[0-9a-f]+//g' | sort | uniq -c
  echo "## memcheck"; timeout 400 compute-sanitizer --tool memcheck python __graft_entry__.py --smoke-only 2>&1 | grep -E "ERROR SUMMARY|smoke ok" ) > gpurun_out/${R}_sanitizer.txt 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_ref_n1.json 2>&1
timeout 300 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/${R}_bench_c5.json 2> gpurun_out/${R}_bench_c5.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_raw.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-api > gpurun_out/${R}_launches.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_bin_rows|k_lwa_cols|k_minmax_bulk|k_scan_epilogue" -s 8 -c 4 -o gpurun_out/prof_${R}_final python bench.py --steps 1 --warmup 3 --no-cpu --no-api > gpurun_out/${R}_final_full.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_${R}_final.ncu-rep 14 > gpurun_out/${R}_final_ncu_full_summary.txt 2>&1
( python scripts/ncu_source_lines.py k_bin_rows 40 gpurun_out/prof_${R}_final.ncu-rep; python scripts/ncu_source_lines.py k_lwa_cols 50 gpurun_out/prof_${R}_final.ncu-rep ) > gpurun_out/${R}_final_source_lines.txt 2>&1
tail -3 gpurun_out/${R}_gpu_tests.txt; cat gpurun_out/${R}_sanitizer.txt; cut -c1-300 gpurun_out/${R}_bench_n1.json
