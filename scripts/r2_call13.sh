# round 2, GPU call 13 (2 GPUs): bench N=2 with the gather inside the timed region, partition independence, streamed run, PCIe ceilings
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "rc=$?" >> gpurun_out/r2_bench_n2.err
timeout 200 $TR scripts/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check_n2.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_multi_gpu_check_n2.txt
timeout 300 $TR scripts/run_c4.py --slices 1024 > gpurun_out/r2_run_c4_n2.json 2> gpurun_out/r2_run_c4_n2.err; echo "rc=$?" >> gpurun_out/r2_run_c4_n2.err
timeout 200 python scripts/run_c4.py --slices 1024 > gpurun_out/r2_run_c4_n1.json 2> gpurun_out/r2_run_c4_n1.err; echo "rc=$?" >> gpurun_out/r2_run_c4_n1.err
timeout 200 $TR scripts/pcie_ceiling.py 512 > gpurun_out/r2_pcie_n2.txt 2>&1
timeout 100 python scripts/pcie_ceiling.py 512 > gpurun_out/r2_pcie_n1.txt 2>&1
