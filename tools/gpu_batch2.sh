#!/bin/bash
set -u
mkdir -p gpurun_out
bash tools/gpu_variants.sh base n2
# query: two CTAs per SM on a 2-deep ring (default) vs one CTA on a 4-deep ring
for st in 2 4; do
  SAF_QUERY_STAGES=$st timeout -s KILL 200 python bench_query.py --rows 12000000 > gpurun_out/q_stages$st.json 2> gpurun_out/q_stages$st.err
  echo "== query stages=$st rc=$?"; tail -c 900 gpurun_out/q_stages$st.json; echo
done
timeout -s KILL 400 python -m pytest tests -q -m gpu > gpurun_out/tests_all.log 2>&1
echo "== all gpu tests rc=$?"; tail -6 gpurun_out/tests_all.log
