# usage: ab_variants.sh v1 v2 ...   (each run with LWA dedupe = match and = tag)
for v in "$@"; do
  for d in m t; do
    XCB200_LWA_DEDUP=$d XCB200_LIB=$PWD/xcontour_b200/libxcb200_$v.so python scripts/time_stages.py 32 16 2>&1 | tail -1 | sed "s/^/[$d] /"
  done
done
