#!/bin/bash
# Everything the round-end driver runs, on one GPU box: pytest -m gpu, the same library tests on the -DPGT_BOUNDS build,
# smoke(), python bench.py.  usage: gpurun --timeout 5400 -- bash tools/run_gpu_regression.sh
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/regress_gpu_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/regress_gpu_tests.log
timeout 1200 bash tools/run_bounds_gpu.sh > gpurun_out/regress_bounds_tests.log 2>&1
echo "bounds rc=$?" >> gpurun_out/regress_bounds_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/regress_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/regress_smoke.log
timeout 1500 python bench.py > gpurun_out/regress_bench_n1.json 2> gpurun_out/regress_bench_n1.err
echo "bench rc=$?" >> gpurun_out/regress_bench_n1.err
tail -n 4 gpurun_out/regress_gpu_tests.log gpurun_out/regress_bounds_tests.log gpurun_out/regress_smoke.log gpurun_out/regress_bench_n1.err
