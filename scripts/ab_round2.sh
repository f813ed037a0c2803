# Round-2 starting point: A/B of the prepared k_lwa_fx variant (XC_FX_OWN=1: own-slot deposits made with
# shared-memory atomics, prefix phase without global loads -- see DESIGN.md §8 and profiles/r1_time_split.txt).
#
#   1. on the CPU (build container):
#        python -c "from xcontour_b200 import build as b; b.build(variant='own', defines=['XC_FX_OWN=1'])"
#   2. on the GPU:   gpurun --timeout 300 -- 'bash scripts/ab_round2.sh'
#
# The variant is held to the full parity suite first (XCB200_LIB selects the library for every test), then
# timed stage by stage against the default build on the benchmark field, a smooth one and a quantised one.
mkdir -p gpurun_out
L=$PWD/xcontour_b200/libxcb200_own.so
( XCB200_LIB=$L timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  for env in "" "XC_NOISE=0" "XC_QUANT=8"; do
    echo "== field: ${env:-benchmark}"
    env $env python scripts/time_stages.py 32 32
    env $env XCB200_LIB=$L python scripts/time_stages.py 32 32
  done ) 2>&1 | grep -v Warning | tee gpurun_out/r2_ab_own.txt
