#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_cli_gpu.py tests/test_sharded_gpu.py tests/test_slide_gpu.py -m gpu -x -q > gpurun_out/r02n_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02n_tests.log
tail -n 25 gpurun_out/r02n_tests.log
