#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_gpu_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02f_gpu_tests.log
timeout 1200 bash tools/run_bounds_gpu.sh > gpurun_out/r02f_bounds_tests.log 2>&1
echo "bounds rc=$?" >> gpurun_out/r02f_bounds_tests.log
timeout 300 python tools/probe_startup.py > gpurun_out/r02f_startup.jsonl 2>&1
timeout 600 python tools/probe_cli_large.py 6e8 0 > gpurun_out/r02f_cli_large.jsonl 2> gpurun_out/r02f_cli_large.err
tail -n 4 gpurun_out/r02f_gpu_tests.log gpurun_out/r02f_bounds_tests.log; cat gpurun_out/r02f_startup.jsonl gpurun_out/r02f_cli_large.jsonl
