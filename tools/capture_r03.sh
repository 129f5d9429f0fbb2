#!/bin/bash
# ncu captures of the kernels that changed late in round 2 (run each command once WITHOUT ncu first)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
python tools/probe_bw.py fused,1e8,1000,1,0,0 het,1e8,1,1,0,0 fst,3e9,50000,10000,512,0 > gpurun_out/r03g_plain.txt 2>&1 || exit 1
timeout 600 $NCU -k regex:k_slide_fused2 -s 2 -c 1 -f -o gpurun_out/r03g_slide_fused2 python tools/probe_bw.py fused,1e8,1000,1,0,0 > gpurun_out/r03g_ncu1.log 2>&1
timeout 600 $NCU -k regex:k_windows_persite4 -s 2 -c 1 -f -o gpurun_out/r03g_persite4 python tools/probe_bw.py het,1e8,1,1,0,0 > gpurun_out/r03g_ncu2.log 2>&1
timeout 600 $NCU -k 'regex:^k_windows$' -s 2 -c 1 -f -o gpurun_out/r03g_windows python tools/probe_bw.py fst,3e9,50000,10000,512,0 > gpurun_out/r03g_ncu3.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --config-steps 2 > gpurun_out/r03g_bench_plain.json 2> gpurun_out/r03g_bench_plain.err || exit 1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r03g_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --config-steps 2 > gpurun_out/r03g_bench_ncu.log 2>&1
ls -la gpurun_out/r03g_*
