#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_paths_gpu.py tests/test_soak_gpu.py tests/test_stats_gpu.py -m gpu -x -q > gpurun_out/r02o_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02o_tests.log
PGT_LIB=$PWD/popgenomicstools_b200/libpgtscan_bounds.so timeout 900 python -m pytest tests/test_paths_gpu.py -m gpu -x -q > gpurun_out/r02o_bounds.log 2>&1
echo "bounds rc=$?" >> gpurun_out/r02o_bounds.log
for NC in 1000 100000 1000000; do
  PROBE_CONTIGS=$NC PGT_TUNE=unittable=1 timeout 300 python tools/probe_bw.py fst,3e8,1000,1000,0,0 fused,3e8,1000,1000,0,0 >> gpurun_out/r02o_probe.log 2>&1
  PROBE_CONTIGS=$NC PGT_TUNE=unittable=2 timeout 300 python tools/probe_bw.py fst,3e8,1000,1000,0,0 fused,3e8,1000,1000,0,0 >> gpurun_out/r02o_probe.log 2>&1
done
tail -n 4 gpurun_out/r02o_tests.log gpurun_out/r02o_bounds.log; cat gpurun_out/r02o_probe.log
