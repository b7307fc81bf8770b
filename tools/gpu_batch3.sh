#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests -x -q -m gpu -k "window_kernels_every_width or sequence_window or segment_table" > gpurun_out/b3_tests.log 2>&1
echo "== tests rc=$? $(tail -1 gpurun_out/b3_tests.log)"
TAG=2gpu_cyclic bash tools/gpu_multi.sh 2
TAG=2gpu_contig bash tools/gpu_multi.sh 2 --slab-layout contiguous --no-query
