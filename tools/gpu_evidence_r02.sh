#!/bin/bash
# Round-2 evidence on one GPU: full GPU test suite, the default bench line and the reference arm, cfg3 on one GPU,
# launch list + --set full captures of the window path and the query GEMM, config-5 cells that fit one GPU.
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/ev_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/ev_tests.log)"
timeout -s KILL 400 python bench.py > gpurun_out/ev_bench.json 2> gpurun_out/ev_bench.err; echo "bench rc=$?"
timeout -s KILL 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ev_ref.json 2> gpurun_out/ev_ref.err; echo "ref rc=$?"
timeout -s KILL 300 python bench.py --window 1 --pool 200 --steps 2 --warmup 1 --no-cpu-baseline --no-query > gpurun_out/ev_bench_window1.json 2> gpurun_out/ev_bench_window1.err; echo "window1 rc=$?"
bash tools/gpu_cfg3_1gpu.sh
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-query"
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 300 --csv --log-file gpurun_out/ev_launches_window.csv $B > gpurun_out/ev_ncu_launches.log 2>&1
echo "launch list rc=$?"
P="python tools/prof_window.py 16 3 1"
for spec in "k3w window_tile_kernel 2" "k2t window_tile_setup 2" "k2 tsdf_update 1" "k1 frame_setup 1"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/ev_$1 -f $P > gpurun_out/ev_ncu_$1.log 2>&1
  echo "ncu $1 rc=$?"
done
Q="python bench_query.py --rows 4000000 --iters 2 --cpu-rows 20000"
$Q > gpurun_out/ev_query_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:query_gemm_tf32_persistent -s 2 -c 2 -o gpurun_out/ev_k4 -f $Q > gpurun_out/ev_ncu_k4.log 2>&1
echo "ncu k4 rc=$?"
timeout -s KILL 200 python bench_query.py --rows 12000000 --iters 3 --cpu-rows 20000 > gpurun_out/ev_query.json 2> gpurun_out/ev_query.err; echo "query rc=$?"
bash tools/gpu_cfg5.sh 1 4 512 4 768 4 1024 2 512 2 768 2 1024
