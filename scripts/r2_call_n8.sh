# round 2, 8-GPU call: the north-star run (8760 slices streamed), bench N=8 with the gather in the timed region, PCIe ceilings
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621"
timeout 240 $TR scripts/run_c4.py --slices 8760 > gpurun_out/r2_run_c4_n8.json 2> gpurun_out/r2_run_c4_n8.err; echo "rc=$?" >> gpurun_out/r2_run_c4_n8.err
timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-api > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "rc=$?" >> gpurun_out/r2_bench_n8.err
timeout 120 $TR scripts/pcie_ceiling.py 256 > gpurun_out/r2_pcie_n8.txt 2>&1
nvidia-smi topo -m > gpurun_out/r2_topo_n8.txt 2>&1
