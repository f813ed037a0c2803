# round 2, GPU call 15 (8 GPUs): bench N=8 with the gather every 5 steps, default NCCL configuration
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641"
timeout 240 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-api > gpurun_out/r2_bench_n8b.json 2> gpurun_out/r2_bench_n8b.err; echo "rc=$?" >> gpurun_out/r2_bench_n8b.err
timeout 200 $TR bench.py --gpus 8 --steps 20 --warmup 3 --no-api --no-cpu --gather-every 1 > gpurun_out/r2_bench_n8c.json 2> gpurun_out/r2_bench_n8c.err; echo "rc=$?" >> gpurun_out/r2_bench_n8c.err
