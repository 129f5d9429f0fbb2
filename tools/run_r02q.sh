#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_slide_gpu.py tests/test_soak_gpu.py tests/test_guard_pages_gpu.py -m gpu -x -q > gpurun_out/r02q_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02q_tests.log
timeout 300 python tools/probe_bw.py fused,1e8,1000,1,0,0 fused,1e8,1000,7,0,0 fused,1e8,600,1,0,0 > gpurun_out/r02q_probe.log 2>&1
PGT_TUNE=fused2=1 timeout 300 python tools/probe_bw.py fused,1e8,1000,1,0,0 fused,1e8,1000,7,0,0 fused,1e8,600,1,0,0 >> gpurun_out/r02q_probe.log 2>&1
tail -n 4 gpurun_out/r02q_tests.log; cat gpurun_out/r02q_probe.log
