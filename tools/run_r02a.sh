#!/bin/bash
# round-2 first GPU call: new-kernel tests, then probes (slide vs unit path at S=1; het ring vs direct; C4 regression)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r02a_box.txt 2>&1
timeout 900 python -m pytest tests/test_slide_gpu.py tests/test_paths_gpu.py tests/test_soak_gpu.py -m gpu -x -q > gpurun_out/r02a_tests1.log 2>&1
echo "tests1 rc=$?" >> gpurun_out/r02a_tests1.log
timeout 600 python tools/probe_bw.py fst,1e8,1000,1,0,0 fst,1e8,1000,7,0,0 fused,1e8,1000,1,0,0 het,1e8,1000,1,0,0 dxy,1e8,1000,1,0,0 fst,3e9,50000,10000,512,0 het,3e9,100000,100000,4096,0 het,3e9,100000,100000,4096,1 het,1e8,100000,100000,4096,0 het,3e9,100000,20000,256,0 fused,1e9,1000,100,0,0 > gpurun_out/r02a_probe.log 2>&1
PGT_TUNE=slide=1 timeout 300 python tools/probe_bw.py fst,1e8,1000,1,0,0 fused,1e8,1000,1,0,0 >> gpurun_out/r02a_probe.log 2>&1
timeout 900 python -m pytest tests/test_stats_gpu.py tests/test_fullscale_gpu.py -m gpu -x -q > gpurun_out/r02a_tests2.log 2>&1
echo "tests2 rc=$?" >> gpurun_out/r02a_tests2.log
tail -3 gpurun_out/r02a_tests1.log gpurun_out/r02a_tests2.log; cat gpurun_out/r02a_probe.log
