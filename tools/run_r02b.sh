#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_slide_gpu.py tests/test_soak_gpu.py -m gpu -x -q > gpurun_out/r02b_tests1.log 2>&1
echo "tests1 rc=$?" >> gpurun_out/r02b_tests1.log
timeout 600 python tools/probe_bw.py fst,1e8,1000,1,0,0 fst,1e8,1000,7,0,0 fused,1e8,1000,1,0,0 het,1e8,1000,1,0,0 dxy,1e8,1000,1,0,0 fst,1e8,256,1,0,0 fst,1e8,1280,16,0,0 > gpurun_out/r02b_probe.log 2>&1
tail -n 3 gpurun_out/r02b_tests1.log; cat gpurun_out/r02b_probe.log
