#!/bin/bash
# First contact with the GPU for a new kernel (run under gpurun): the whole GPU test suite, then the bench per
# window-kernel variant, then the default bench line and a launch list.
set -u
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/ab_host.txt; nproc >> gpurun_out/ab_host.txt; nvidia-smi -L >> gpurun_out/ab_host.txt
timeout -s KILL 300 python -m pytest tests -x -q -m gpu -k "sequence or window or round2" > gpurun_out/ab_tests.log 2>&1
trc=$?
echo "window tests rc=$trc" | tee -a gpurun_out/ab_tests.log
tail -5 gpurun_out/ab_tests.log
variants="1 2"
if [ $trc -eq 137 ]; then variants="1"; fi     # a hang in the new kernel: time the old one and stop
for v in $variants; do
  SAF_K3W_VARIANT=$v timeout -s KILL 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-query \
      > gpurun_out/ab_variant$v.json 2> gpurun_out/ab_variant$v.err
  echo "variant $v rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_variant$v.json"))
    r = d["roofline"]
    print("variant $v: value %.3e ms/step %.3f k3w_us %.1f k2_us %.1f k1_us %.1f upd/launch %.0f union %.0f frac %.3f" % (d["value"], d["ms_per_step"], r["avg_launch_us"], r["k2_avg_us"], r["k1_avg_us"], r["avg_updates_per_launch"], r["avg_union_rows_per_launch"], r["frac"]))
except Exception as e:
    print("variant $v: no line", e)
PY
done
if [ $trc -eq 137 ]; then exit 0; fi
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 > gpurun_out/ab_full.json 2> gpurun_out/ab_full.err
echo "full bench rc=$?"; tail -c 1500 gpurun_out/ab_full.json; tail -5 gpurun_out/ab_full.err
timeout -s KILL 400 python -m pytest tests -q -m gpu > gpurun_out/ab_tests_all.log 2>&1
echo "all gpu tests rc=$?"; tail -8 gpurun_out/ab_tests_all.log
