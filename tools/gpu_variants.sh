#!/bin/bash
# A/B of experimental builds of the library (python -m spatially_aware_ai_b200.build --variant NAME -D...): parity
# tests of the window path, then the bench, per variant.  usage: tools/gpu_variants.sh base p2 f2 ...
set -u
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = "base" ]; then unset SAF_LIB_PATH; else export SAF_LIB_PATH=$PWD/spatially_aware_ai_b200/libsaf_b200_$v.so; fi
  if [ -n "${SKIP_TESTS:-}" ]; then echo "== $v: tests skipped"; else
  timeout -s KILL 240 python -m pytest tests -x -q -m gpu -k "window_kernels_every_width or sequence_window or sensor_format or segment_table or block_cyclic" \
      > gpurun_out/var_${v}_tests.log 2>&1
  trc=$?
  echo "== $v: tests rc=$trc $(tail -1 gpurun_out/var_${v}_tests.log)"
  if [ $trc -ne 0 ]; then tail -15 gpurun_out/var_${v}_tests.log; continue; fi
  fi
  timeout -s KILL 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-query \
      > gpurun_out/var_${v}.json 2> gpurun_out/var_${v}.err
  echo "== $v: bench rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/var_${v}.json"))
    r = d["roofline"]
    print("   $v: value %.3e ms/step %.3f k3w_us %.1f ns/upd %.3f k2_us %.1f k1_us %.1f upd/launch %.0f union %.0f frac %.3f" % (d["value"], d["ms_per_step"], r["avg_launch_us"], r["ns_per_update"], r["k2_avg_us"], r["k1_avg_us"], r["avg_updates_per_launch"], r["avg_union_rows_per_launch"], r["frac"]))
except Exception as e:
    print("   $v: no line", e)
PY
done
