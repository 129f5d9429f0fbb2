#!/bin/bash
# 8-GPU call: bench at N=8 (what the driver's scaling run does), topology, per-GPU H2D
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02g_topo.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02g_bench_n8.json 2> gpurun_out/r02g_bench_n8.err
echo "bench rc=$?" >> gpurun_out/r02g_bench_n8.err
tail -n 6 gpurun_out/r02g_bench_n8.err; grep '^{' gpurun_out/r02g_bench_n8.json | head -c 2500
