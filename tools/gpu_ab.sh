#!/bin/bash
# A/B of the window feature kernels on one B200 (run under gpurun): parity tests first, then the bench per variant.
set -u
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "sequence or window" > gpurun_out/ab_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/ab_tests.log
tail -5 gpurun_out/ab_tests.log
for v in 1 2; do
  SAF_K3W_VARIANT=$v timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e \
      > gpurun_out/ab_variant$v.json 2> gpurun_out/ab_variant$v.err
  echo "variant $v rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_variant$v.json"))
    r = d["roofline"]
    print("variant $v: value %.3e ms/step %.3f k3w_us %.1f k2_us %.1f k1_us %.1f upd/launch %.0f" % (d["value"], d["ms_per_step"], r["avg_launch_us"], r["k2_avg_us"], r["k1_avg_us"], r["avg_updates_per_launch"]))
except Exception as e:
    print("variant $v: no line", e)
PY
done
