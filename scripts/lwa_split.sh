# Where the time of the two heavy kernels goes (1 GPU).  The variant libraries are built on the CPU first:
#   python -c "from xcontour_b200 import build as b; b.build(variant='lwa1', defines=['XC_FX_EXP=1']); \
#              b.build(variant='lwa2', defines=['XC_FX_EXP=2']); b.build(variant='hk1', defines=['XC_HKX_EXP=1'])"
# Their results are wrong by construction; only the stage times are read.
mkdir -p gpurun_out
L=xcontour_b200
( python scripts/time_stages.py 32 32
  XCB200_LIB=$L/libxcb200_lwa1.so python scripts/time_stages.py 32 32
  XCB200_LIB=$L/libxcb200_lwa2.so python scripts/time_stages.py 32 32
  XC_NO_LWA=1 XCB200_LIB=$L/libxcb200_hk1.so python scripts/time_stages.py 32 32 ) 2>&1 | grep -v Warning > gpurun_out/r1_split.txt
cat gpurun_out/r1_split.txt
