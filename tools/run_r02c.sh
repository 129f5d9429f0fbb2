#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_slide_gpu.py tests/test_soak_gpu.py -m gpu -x -q > gpurun_out/r02c_tests1.log 2>&1
echo "tests1 rc=$?" >> gpurun_out/r02c_tests1.log
timeout 600 python tools/probe_bw.py fst,1e8,1000,1,0,0 fst,1e8,1000,7,0,0 fused,1e8,1000,1,0,0 het,1e8,1000,1,0,0 fst,1e8,256,1,0,0 > gpurun_out/r02c_probe.log 2>&1
timeout 1500 python bench.py > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r02c_bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_slide -c 1 -o gpurun_out/r02c_prof_slide python tools/probe_bw.py fst,1e8,1000,1,0,0 > gpurun_out/r02c_ncu_slide.log 2>&1
tail -n 3 gpurun_out/r02c_tests1.log; cat gpurun_out/r02c_probe.log; tail -n 5 gpurun_out/r02c_bench_n1.err; head -c 6000 gpurun_out/r02c_bench_n1.json
