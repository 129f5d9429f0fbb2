#!/bin/bash
# 2-GPU call: sharded tests over real devices, bench at N=2 (torchrun), quick probes at N=1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02d_topo.txt 2>&1
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_stats_gpu.py tests/test_paths_gpu.py -m gpu -x -q > gpurun_out/r02d_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02d_tests.log
timeout 300 python tools/probe_bw.py het,3e9,100000,100000,4096,0 het,3e9,100000,100000,4096,1 het,1e8,1,1,0,0 fst,1e8,1,1,0,0 > gpurun_out/r02d_probe.log 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02d_bench_n2.json 2> gpurun_out/r02d_bench_n2.err
echo "bench rc=$?" >> gpurun_out/r02d_bench_n2.err
tail -n 3 gpurun_out/r02d_tests.log; cat gpurun_out/r02d_probe.log; tail -n 5 gpurun_out/r02d_bench_n2.err; head -c 3000 gpurun_out/r02d_bench_n2.json
