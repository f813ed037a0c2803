# one --set full capture of the two heavy kernels (1 launch each, after warm-up)
export XCB200_SUB_BATCH=16
TAG=${1:-v1}
ncu --set full --clock-control none --import-source on -k regex:"^k_hist$|k_lwa_fast" -s 6 -c 2 -o gpurun_out/prof_r1_$TAG python bench.py --steps 1 --warmup 3 --batch 32 --no-cpu > gpurun_out/b_ncu_$TAG.log 2>&1
tail -1 gpurun_out/b_ncu_$TAG.log | cut -c1-300
