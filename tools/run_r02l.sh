#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02l_gpu_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02l_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02l_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r02l_smoke.log
timeout 1500 python bench.py > gpurun_out/r02l_bench_n1.json 2> gpurun_out/r02l_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r02l_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02l_bench_ref.json 2> gpurun_out/r02l_bench_ref.err
tail -n 4 gpurun_out/r02l_gpu_tests.log gpurun_out/r02l_smoke.log gpurun_out/r02l_bench_n1.err; head -c 1500 gpurun_out/r02l_bench_ref.json
