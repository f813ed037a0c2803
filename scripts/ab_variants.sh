# usage: ab_variants.sh v1 v2 ...   (per-stage times of experimental builds; XC_NOISE / XC_QUANT select the field)
for v in "$@"; do
    XCB200_LIB=$PWD/xcontour_b200/libxcb200_$v.so python scripts/time_stages.py 32 16 2>&1 | tail -1
done
