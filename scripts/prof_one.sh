# usage: prof_one.sh <kernel-regex> <tag> [skip]
export XCB200_SUB_BATCH=16
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${3:-3} -c 1 -o gpurun_out/prof_r1_$2 python scripts/time_stages.py 32 16 > gpurun_out/p_$2.log 2>&1
tail -1 gpurun_out/p_$2.log | cut -c1-200
