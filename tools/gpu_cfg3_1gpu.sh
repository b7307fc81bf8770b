#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 500 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_cfg3_1gpu.json 2> gpurun_out/b_cfg3_1gpu.err
echo "rc=$?"; grep "\[bench\]" gpurun_out/b_cfg3_1gpu.err | cut -c1-200
python - <<PY
import json
d = json.load(open("gpurun_out/b_cfg3_1gpu.json"))
print("cfg3 N=1 value %.3e ms/step %.3f e2e %.3e query %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["query"]))
print(d["roofline"]["avg_launch_us"], d["roofline"]["ns_per_update"], d["updates_per_frame"], d["updates_per_union_row"])
PY
