set -x
export XCB200_SUB_BATCH=16
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_v0.csv python bench.py --steps 1 --warmup 3 --batch 32 --no-cpu > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_hist|k_lwa_fast|k_minmax" -s 9 -c 3 -o gpurun_out/prof_r1_v0 python bench.py --steps 1 --warmup 3 --batch 32 --no-cpu > gpurun_out/b_ncu2.log 2>&1
tail -2 gpurun_out/b_ncu2.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
