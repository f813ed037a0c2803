#!/bin/bash
# N = 2: bench with the default per-step gather (default NCCL configuration) + bit-identity of the 2-rank run
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
  bench.py --gpus 2 --steps 10 --warmup 3 --no-api > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 \
  scripts/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check_n2.txt 2>&1
cut -c1-200 gpurun_out/r2_bench_n2.json; tail -3 gpurun_out/r2_multi_gpu_check_n2.txt
