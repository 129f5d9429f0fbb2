#!/bin/bash
# 8-GPU call: where is the ceiling of host -> device bandwidth on this box?  raw pinned copies vs pgt_scan_sharded, 1/2/4/8 GPUs
mkdir -p gpurun_out
L=gpurun_out/r02m_e2e_8gpu.log
: > $L
python tools/probe_e2e_multi.py 2e9 0,1,2,3,4,5,6,7 raw >> $L 2>&1
python tools/probe_e2e_multi.py 2e9 0,1,2,3 raw >> $L 2>&1
python tools/probe_e2e_multi.py 2e9 0,2,4,6 raw >> $L 2>&1
python tools/probe_e2e_multi.py 2e9 0,1,2,3,4,5,6,7 >> $L 2>&1
python tools/probe_e2e_multi.py 2e9 0,1,2,3 >> $L 2>&1
grep -v Warning $L
