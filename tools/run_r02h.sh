#!/bin/bash
# ncu evidence of round 2: launch list of the bench command, full captures of the dominant kernels
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --configs C3-1/1,C3-100k,C5-S1 --config-steps 2"
timeout 600 $B > gpurun_out/r02h_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02h_launches.csv $B > gpurun_out/r02h_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_units_tiled -s 3 -c 1 -o gpurun_out/r02h_prof_tiled python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r02h_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_units_het_vec -s 1 -c 1 -o gpurun_out/r02h_prof_hetvec python tools/probe_bw.py het,3e9,100000,100000,4096,0 > gpurun_out/r02h_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_windows_persite -s 1 -c 1 -o gpurun_out/r02h_prof_persite python tools/probe_bw.py het,1e8,1,1,0,0 > gpurun_out/r02h_ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_slide -s 1 -c 1 -o gpurun_out/r02h_prof_slide_fused python tools/probe_bw.py fused,1e8,1000,1,0,0 > gpurun_out/r02h_ncu5.log 2>&1
tail -n 2 gpurun_out/r02h_ncu*.log; wc -l gpurun_out/r02h_launches.csv
