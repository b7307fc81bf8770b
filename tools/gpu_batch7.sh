#!/bin/bash
set -u
mkdir -p gpurun_out
bash tools/gpu_variants.sh base p1
timeout -s KILL 200 python bench.py --emulate-world 8 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-query > gpurun_out/b7_emu8.json 2> gpurun_out/b7_emu8.err
python - <<PY
import json
d = json.load(open("gpurun_out/b7_emu8.json")); r = d["roofline"]
print("   emu8: ms/step %.3f value %.3e k3w_us %.1f k2_us %.1f k1_us %.1f" % (d["ms_per_step"], d["value"], r["avg_launch_us"], r["k2_avg_us"], r["k1_avg_us"]))
PY
