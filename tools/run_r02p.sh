#!/bin/bash
mkdir -p gpurun_out
PGT_TUNE=unittable=1 timeout 300 python tools/probe_bw.py fst,3e9,50000,10000,512,0 fused,1e9,1000,100,0,0 > gpurun_out/r02p_probe.log 2>&1
PGT_TUNE=unittable=2 timeout 300 python tools/probe_bw.py fst,3e9,50000,10000,512,0 fused,1e9,1000,100,0,0 >> gpurun_out/r02p_probe.log 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02p_bench_n4.json 2> gpurun_out/r02p_bench_n4.err
echo "bench rc=$?" >> gpurun_out/r02p_bench_n4.err
cat gpurun_out/r02p_probe.log; tail -n 3 gpurun_out/r02p_bench_n4.err; grep '^{' gpurun_out/r02p_bench_n4.json | head -c 1200
