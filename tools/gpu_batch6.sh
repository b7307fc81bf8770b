#!/bin/bash
set -u
mkdir -p gpurun_out
for r in 0 10 18 28 40; do
  SAF_K3W_RESERVE_SMS=$r timeout -s KILL 200 python bench.py --emulate-world 8 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-query > gpurun_out/b6_emu8_r$r.json 2> gpurun_out/b6_emu8_r$r.err
  python - <<PY
import json
d = json.load(open("gpurun_out/b6_emu8_r$r.json")); r = d["roofline"]
print("   emu8 reserve $r: ms/step %.3f value %.3e k3w_us %.1f k2_us %.1f k1_us %.1f" % (d["ms_per_step"], d["value"], r["avg_launch_us"], r["k2_avg_us"], r["k1_avg_us"]))
PY
done
for r in 0 18; do
  SAF_K3W_RESERVE_SMS=$r timeout -s KILL 200 python bench.py --emulate-world 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-query > gpurun_out/b6_emu2_r$r.json 2> gpurun_out/b6_emu2_r$r.err
  python - <<PY
import json
d = json.load(open("gpurun_out/b6_emu2_r$r.json")); r = d["roofline"]
print("   emu2 reserve $r: ms/step %.3f value %.3e k3w_us %.1f k2_us %.1f k1_us %.1f" % (d["ms_per_step"], d["value"], r["avg_launch_us"], r["k2_avg_us"], r["k1_avg_us"]))
PY
done
