# round 2, GPU call 9: full GPU suite with durations after the plane-zeroing fix
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -q --durations=12 2>&1 | tail -40
  python scripts/time_stages.py 32 32 ) 2>&1 | grep -v Warning | tee gpurun_out/r2_call9.txt
