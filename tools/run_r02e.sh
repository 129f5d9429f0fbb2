#!/bin/bash
mkdir -p gpurun_out
df -h /dev/shm /tmp > gpurun_out/r02e_df.txt 2>&1; free -g >> gpurun_out/r02e_df.txt; nproc >> gpurun_out/r02e_df.txt
timeout 900 python -m pytest tests/test_cli_gpu.py tests/test_extreme_cli_gpu.py tests/test_binding_gpu.py -m gpu -x -q > gpurun_out/r02e_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02e_tests.log
timeout 1200 python tools/probe_cli_large.py 6e8 2.4e8 > gpurun_out/r02e_cli_large.jsonl 2> gpurun_out/r02e_cli_large.err
tail -n 4 gpurun_out/r02e_tests.log; cat gpurun_out/r02e_df.txt; cat gpurun_out/r02e_cli_large.jsonl; tail -n 5 gpurun_out/r02e_cli_large.err
