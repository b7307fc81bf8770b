#!/bin/bash
# ncu --set full capture of the window feature kernel (run under gpurun, one GPU)
set -u
mkdir -p gpurun_out
python tools/prof_window.py 16 3 1 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:window_tile -s 1 -c 2 \
    -o gpurun_out/r02_k3w_tile_p2 -f python tools/prof_window.py 16 3 1 > gpurun_out/prof_ncu.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/prof_plain.log; tail -3 gpurun_out/prof_ncu.log
