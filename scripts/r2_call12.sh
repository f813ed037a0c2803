# round 2, GPU call 12: after the planner fix -- config 5 probe, the tests that were cut off, LWA SEG=32 A/B, smooth-field binning
mkdir -p gpurun_out
timeout 100 python -u scripts/c5_probe.py > gpurun_out/r2_c5probe_fixed.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_c5probe_fixed.txt
timeout 400 python -u -m pytest tests/test_gpu_bench_configs.py -m gpu -q -k "gradient or gather or f32 or c5 or row_march" --durations=6 > gpurun_out/r2_tests12.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_tests12.txt
( python scripts/time_stages.py 32 32
  XCB200_LWA_SEG=32 python scripts/time_stages.py 32 32
  XC_NOISE=0 python scripts/time_stages.py 32 32
  XCB200_LWA_SEG=32 timeout 120 python __graft_entry__.py --smoke-only 2>&1 | tail -1
  timeout 200 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 ) > gpurun_out/r2_call12.txt 2>&1
