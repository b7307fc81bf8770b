#!/bin/bash
# query path: parity tests, then bench_query.py for each "NAME ENV=..." argument
set -u
mkdir -p gpurun_out
if [ -z "${SKIP_TESTS:-}" ]; then
timeout -s KILL 300 python -m pytest tests -x -q -m gpu -k "query or topk or surgery or labels or presence or segment_labels" > gpurun_out/query_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/query_tests.log)"; grep -E "^(FAILED|ERROR)" gpurun_out/query_tests.log | head
fi
for spec in "$@"; do
  set -- $spec; name=$1; shift
  env "$@" timeout -s KILL 150 python bench_query.py --rows ${ROWS:-12000000} --iters 3 --cpu-rows 20000 > gpurun_out/query_${name}.json 2> gpurun_out/query_${name}.err
  echo "== $name rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/query_${name}.json"))
    print("   ${name}: scores %.2f ms (%.0f GB/s read, %.0f r+w, %.0f TFLOP/s)  topk %.2f ms (%.0f GB/s)  err %.2e" % (d["scores_ms"], d["read_GBps"], d["read_plus_write_GBps"], d["tflops"], d["topk_ms"], d["topk_read_GBps"], d["max_abs_err_vs_oracle"]))
except Exception as e:
    print("   ${name}: no line", e); print(open("gpurun_out/query_${name}.err").read()[-1500:])
PY
done
