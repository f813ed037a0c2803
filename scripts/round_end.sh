# Round-end evidence run (1 GPU): racecheck, launch list, full ncu capture of the three hot kernels, bench.
mkdir -p gpurun_out
( echo "## default (exact integer accumulators on shared-memory atomics in k_lwa_fx / k_hist_keff)"; timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python __graft_entry__.py --smoke 2>&1 | grep -E "Race reported|RACECHECK SUMMARY|smoke ok" | sed -E 's/\+0x[0-9a-f]+//g' | sort | uniq -c
  echo "## XCB200_LWA_FX=0 XCB200_HIST_FX=0 XCB200_LWA_DEDUP=m (fp64 read-modify-write kernels, MATCH.ANY peel everywhere)"; XCB200_LWA_FX=0 XCB200_HIST_FX=0 XCB200_LWA_DEDUP=m timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python __graft_entry__.py --smoke 2>&1 | grep -E "Race reported|RACECHECK SUMMARY|smoke ok" | sort | uniq -c
  echo "## memcheck"; timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py --smoke 2>&1 | grep -E "ERROR SUMMARY|smoke ok" ) > gpurun_out/r1_sanitizer.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_final_launches_raw.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1_final_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_hist_keff|^k_lwa_fx$|k_minmax_partial|k_scan_epilogue" -s 8 -c 4 -o gpurun_out/prof_r1_final python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r1_final_full.log 2>&1
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r1_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/r1_bench_ref_n1.json
python scripts/c5_check.py 2>&1 | tail -1 > gpurun_out/r1_c5.txt
cat gpurun_out/r1_sanitizer.txt; cut -c1-400 gpurun_out/r1_bench_n1.json; cat gpurun_out/r1_c5.txt
