# Round-2 starting point: A/B of the variants prepared (and not yet timed) at the end of round 1.  Every one is a
# compile-time switch whose default leaves the round-1 SASS untouched (profiles/README.md):
#   own    -DXC_FX_OWN=1     k_lwa_fx: own-slot deposits with shared-memory atomics, prefix walk without global loads
#   lut4k  -DXC_FX_LUT=4096  k_lwa_fx: 4x finer LUT over Q (the closing bisection is 9 % of the kernel's instructions)
#   lean   -DXC_HKX_LEAN=1   k_hist_keff: funnel-shift decomposition, out-of-line truncation, 64-bit-add carries,
#                            no division in the cell loop (hkx_add is 55 % of the kernel's instructions)
#   rowcnt -DXC_HKX_ROWCNT=1 -DXC_HKX_LEAN=1   k_hist_keff: area as exact integer cell counts per (row, bin) where dA is
#                            constant along the row (one ATOMS.ADD per cell), windowed add only for |grad q|^2 dA
#   fxlean -DXC_FX_LEAN=1    k_lwa_fx: carries of the 64-bit atomic adds from 64-bit integer adds
#   own8   -DXC_FX_OWN=1 -DXC_FX_TC8=1   the own-deposit kernel on 8-column tiles, two CTAs per SM
#   pair   -DXC_FX_OWN=1 -DXC_FX_PAIR=1  own-deposit kernel with lo/hi words adjacent: LDS.64 in the prefix passes
#   all    own + pair + lut4k + lean + rowcnt + fxlean together
#
#   1. on the CPU (build container):
#        python scripts/build_variants.py own:XC_FX_OWN=1 lut4k:XC_FX_LUT=4096 lean:XC_HKX_LEAN=1 fxlean:XC_FX_LEAN=1 rowcnt:XC_HKX_ROWCNT=1,XC_HKX_LEAN=1 own8:XC_FX_OWN=1,XC_FX_TC8=1 pair:XC_FX_OWN=1,XC_FX_PAIR=1 \
#               all:XC_FX_OWN=1,XC_FX_PAIR=1,XC_FX_LUT=4096,XC_HKX_LEAN=1,XC_HKX_ROWCNT=1,XC_FX_LEAN=1
#   2. on the GPU:   gpurun --timeout 600 -- 'bash scripts/ab_round2.sh'
#
# Each variant is first held to the parity tests that touch its kernel (XCB200_LIB selects the library for every
# test; run the whole suite on the winner before flipping a default), then timed stage by stage against the default
# build on the benchmark field, a smooth one and a quantised one.
mkdir -p gpurun_out
D=$PWD/xcontour_b200
( for v in own lut4k lean rowcnt fxlean own8 pair all; do
    echo "== parity, variant $v"
    XCB200_LIB=$D/libxcb200_$v.so timeout 300 python -m pytest tests -m gpu -x -q \
        -k "lwa or lape or fused or workflow or reference_fixtures or full_size or cdf or hist or keff or accumulators or smoke" 2>&1 | tail -2
  done
  for env in "" "XC_NOISE=0" "XC_QUANT=8"; do
    echo "== field: ${env:-benchmark}"
    env $env python scripts/time_stages.py 32 32
    for v in own lut4k lean rowcnt fxlean own8 pair all; do env $env XCB200_LIB=$D/libxcb200_$v.so python scripts/time_stages.py 32 32; done
  done ) 2>&1 | grep -v Warning | tee gpurun_out/r2_ab.txt
