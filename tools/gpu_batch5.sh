#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -x -q -m gpu -k "window or sequence or topk or round2 or query" > gpurun_out/b5_tests.log 2>&1
echo "== tests rc=$? $(tail -1 gpurun_out/b5_tests.log)"; grep -E "FAILED|Error" gpurun_out/b5_tests.log | head
timeout -s KILL 300 python bench.py --emulate-world 8 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-query > gpurun_out/b5_emu8.json 2> gpurun_out/b5_emu8.err
echo "== emulate-world 8 rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/b5_emu8.json")); r = d["roofline"]
print("   emu8: ms/step %.3f value %.3e k3w_us %.1f k2_us %.1f k1_us %.1f upd/launch %.0f union %.0f ns/upd %.3f" % (d["ms_per_step"], d["value"], r["avg_launch_us"], r["k2_avg_us"], r["k1_avg_us"], r["avg_updates_per_launch"], r["avg_union_rows_per_launch"], r["ns_per_update"]))
PY
timeout -s KILL 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b5_cfg3.json 2> gpurun_out/b5_cfg3.err
echo "== cfg3 rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/b5_cfg3.json")); print("   cfg3 query", d["query"]["ms"], d["query"]["read_gbs"], "value %.3e" % d["value"])
PY
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b5_ref.json 2> gpurun_out/b5_ref.err
echo "== reference arm rc=$?"; cut -c1-400 gpurun_out/b5_ref.json; tail -3 gpurun_out/b5_ref.err
