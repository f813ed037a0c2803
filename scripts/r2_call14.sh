# round 2, GPU call 14 (2 GPUs): bench N=2 with the gather every 5 steps; sub-batch experiments on one GPU
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631"
timeout 200 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-api --no-cpu > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err; echo "rc=$?" >> gpurun_out/r2_bench_n2b.err
timeout 200 $TR bench.py --gpus 2 --steps 7 --warmup 3 --no-api --no-cpu > gpurun_out/r2_bench_n2c.json 2> gpurun_out/r2_bench_n2c.err; echo "rc=$?" >> gpurun_out/r2_bench_n2c.err
for sb in 8 16; do timeout 120 python bench.py --steps 10 --warmup 3 --no-api --no-cpu --sub-batch $sb > gpurun_out/r2_bench_sub$sb.json 2>&1; done
