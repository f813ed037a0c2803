# round 2, GPU call 8: full GPU suite with durations, bench N=1 (ours + reference arm)
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -25
  python scripts/time_stages.py 32 32
  timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 3000 gpurun_out/r2_bench_n1.json; tail -3 gpurun_out/r2_bench_n1.err
  timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_ref_n1.json 2>&1; tail -c 600 gpurun_out/r2_bench_ref_n1.json ) 2>&1 | grep -v Warning | tee gpurun_out/r2_call8.txt
