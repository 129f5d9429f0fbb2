#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02i_e2e_multi.log
: > $L
python tools/probe_e2e_multi.py 1e9 0 >> $L 2>&1
python tools/probe_e2e_multi.py 1e9 1 >> $L 2>&1
python tools/probe_e2e_multi.py 1e9 0,1 >> $L 2>&1
python tools/probe_e2e_multi.py 1e9 0,1 raw >> $L 2>&1
echo "-- two processes side by side, one GPU each" >> $L
python tools/probe_e2e_multi.py 5e8 0 >> $L 2>&1 &
python tools/probe_e2e_multi.py 5e8 1 >> $L 2>&1
wait
echo "-- two processes, raw" >> $L
python tools/probe_e2e_multi.py 5e8 0 raw >> $L 2>&1 &
python tools/probe_e2e_multi.py 5e8 1 raw >> $L 2>&1
wait
grep -v Warning $L
