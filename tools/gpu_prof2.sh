#!/bin/bash
# --set full captures of the window path's kernels (K3W tile kernel, K2T, K2, K1) on cfg2
set -u
mkdir -p gpurun_out
P="python tools/prof_window.py 16 3 1"
$P > gpurun_out/prof2_plain.log 2>&1 || { tail -5 gpurun_out/prof2_plain.log; exit 1; }
tail -2 gpurun_out/prof2_plain.log
for spec in "k3w window_tile_kernel 2" "k2t window_tile_setup 2" "k2 tsdf_update 1" "k1 frame_setup 1"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/${TAG:-r02b}_$1 -f $P > gpurun_out/prof2_$1.log 2>&1
  echo "$1 rc=$?"
done
