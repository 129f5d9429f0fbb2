#!/bin/bash
mkdir -p gpurun_out
PGT_LIB=$PWD/popgenomicstools_b200/libpgtscan_bounds.so timeout 900 python -m pytest tests/test_slide_gpu.py tests/test_soak_gpu.py -m gpu -x -q > gpurun_out/r02j_bounds.log 2>&1
echo "bounds rc=$?" >> gpurun_out/r02j_bounds.log
timeout 1200 python -m pytest tests/test_slide_gpu.py tests/test_soak_gpu.py tests/test_guard_pages_gpu.py tests/test_sharded_gpu.py tests/test_stats_gpu.py -m gpu -x -q > gpurun_out/r02j_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02j_tests.log
timeout 600 python tools/probe_bw.py fst,1e8,1000,1,0,0 fst,1e8,1000,7,0,0 fused,1e8,1000,1,0,0 het,1e8,1000,1,0,0 dxy,1e8,1000,1,0,0 fst,1e8,256,1,0,0 fst,1e8,500,1,0,0 fst,1e8,100,1,0,0 fst,1e8,64,3,0,0 > gpurun_out/r02j_probe.log 2>&1
PGT_TUNE=slide=1 timeout 300 python tools/probe_bw.py fst,1e8,100,1,0,0 fst,1e8,256,1,0,0 >> gpurun_out/r02j_probe.log 2>&1
tail -n 5 gpurun_out/r02j_bounds.log gpurun_out/r02j_tests.log; cat gpurun_out/r02j_probe.log
