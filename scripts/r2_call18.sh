#!/bin/bash
# N = 4: all-gather cadence, same step count, two repetitions each
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for rep in 1 2; do for g in 1 5; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29600+rep*10+g)) \
  bench.py --gpus 4 --steps 10 --warmup 3 --no-api --no-cpu --gather-every $g > gpurun_out/r2_n4_g${g}_rep${rep}.json 2> gpurun_out/r2_n4_g${g}_rep${rep}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_n4_g${g}_rep${rep}.json").read().strip().splitlines()[-1])
print("G=${g} rep=${rep}", d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
done; done
