#!/bin/bash
# Quick check of the window path: parity tests, then kernel-only bench lines for each "NAME=ENV..." argument.
# usage: tools/gpu_quick.sh base "nosetup SAF_TILE_SETUP=0" ...
set -u
mkdir -p gpurun_out
if [ -z "${SKIP_TESTS:-}" ]; then
timeout -s KILL 400 python -m pytest tests -x -q -m gpu -k "${TESTS:-window or sequence or sensor_format or segment_table or block_cyclic or sheared or slab or checkpoint or resume}" > gpurun_out/quick_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/quick_tests.log)"; grep -E "^(FAILED|ERROR)" gpurun_out/quick_tests.log | head
fi
for spec in "$@"; do
  set -- $spec; name=$1; shift
  env "$@" timeout -s KILL 200 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-e2e --no-query ${BENCH_ARGS:-} \
      > gpurun_out/quick_${name}.json 2> gpurun_out/quick_${name}.err
  echo "== $name: bench rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/quick_${name}.json"))
    r = d["roofline"]
    print("   ${name}: value %.3e ms/step %.3f k3w_us %.1f ns/upd %.3f k2_us %.1f k1_us %.1f upd/launch %.0f union %.0f launches %s" % (d["value"], d["ms_per_step"], r["avg_launch_us"], r["ns_per_update"], r["k2_avg_us"], r["k1_avg_us"], r["avg_updates_per_launch"], r["avg_union_rows_per_launch"], d.get("gpu_launches")))
except Exception as e:
    print("   ${name}: no line", e)
    print(open("gpurun_out/quick_${name}.err").read()[-1500:])
PY
done
