#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -x -q -m gpu -k "sheared or mostly_unobserved or block_cyclic or window_kernels_every_width or topk" > gpurun_out/b4_tests.log 2>&1
echo "== tests rc=$? $(tail -1 gpurun_out/b4_tests.log)"; grep -E "FAILED|Error" gpurun_out/b4_tests.log | head
TAG=2gpu_sheared bash tools/gpu_multi.sh 2
