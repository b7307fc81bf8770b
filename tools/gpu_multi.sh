#!/bin/bash
# Multi-GPU bench lines (run under gpurun --gpus N).  usage: tools/gpu_multi.sh N [extra bench args...]
set -u
N=$1; shift
mkdir -p gpurun_out
tag=${TAG:-${N}gpu}
timeout -s KILL ${LIMIT:-420} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps ${STEPS:-10} --warmup ${WARMUP:-3} "$@" > gpurun_out/b_${tag}.json 2> gpurun_out/b_${tag}.err
echo "rc=$?"
grep "\[bench\]" gpurun_out/b_${tag}.err | cut -c1-220
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/b_${tag}.json"))
    print("N=%d value %.3e ms/step %.3f imbalance %.3f per-rank ms %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["rank_imbalance"], ["%.2f" % x for x in d["per_rank_ms_per_step"]]))
    print("  e2e", d["e2e"] and {k: d["e2e"][k] for k in ("value", "frames_per_s", "h2d_bytes_per_step", "host_format")})
    print("  query", d["query"] and {k: d["query"].get(k) for k in ("rows", "ms", "read_gbs", "error")})
    print("  config", d["config"]["workload"][:100], "|", d["config"]["parallelism"])
except Exception as e:
    print("no line", e); import subprocess; print(open("gpurun_out/b_${tag}.err").read()[-3000:])
PY
