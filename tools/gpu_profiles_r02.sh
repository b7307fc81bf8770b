#!/bin/bash
# Round-2 evidence: launch list of the bench's window path, --set full captures of K3W / K2 / K4.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-query"
$B > gpurun_out/r02_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r02_launches_window.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
P="python tools/prof_window.py 16 3 1"
$P > gpurun_out/r02_plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:window_tile -s 1 -c 2 -o gpurun_out/r02_k3w_final -f $P > gpurun_out/r02_ncu_k3w.log 2>&1
echo "k3w rc=$?"
ncu --set full --clock-control none --import-source on -k regex:tsdf_update -s 1 -c 1 -o gpurun_out/r02_k2w_final -f $P > gpurun_out/r02_ncu_k2.log 2>&1
echo "k2 rc=$?"
Q="python bench_query.py --rows 4000000 --iters 2 --cpu-rows 20000"
$Q > gpurun_out/r02_plain_query.log 2>&1 && \
ncu --set full --clock-control none -k regex:query_gemm_tf32 -s 2 -c 2 -o gpurun_out/r02_k4_final -f $Q > gpurun_out/r02_ncu_k4.log 2>&1
echo "k4 rc=$?"
tail -3 gpurun_out/r02_plain_prof.log
