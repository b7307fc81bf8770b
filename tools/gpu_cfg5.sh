#!/bin/bash
# BASELINE config 5 cells: 500 new frames into a grid that already holds 60, feature dim x voxel size.
# usage: tools/gpu_cfg5.sh <n_gpus> <voxel_cm> <C> [<voxel_cm> <C> ...]
set -u
N=$1; shift
mkdir -p gpurun_out
while [ $# -ge 2 ]; do
  vs=$1; C=$2; shift 2
  out=gpurun_out/cfg5_C${C}_${vs}cm_${N}gpu
  args="--workload cfg5 --voxel-size 0.0$vs --feature-dim $C --frames-per-step 100 --warmup 1 --steps 5 --pool 560 --no-cpu-baseline --no-query"
  if [ "$N" = "1" ]; then
    timeout -s KILL 400 python bench.py $args > $out.json 2> $out.err
  else
    timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N $args > $out.json 2> $out.err
  fi
  echo "== cfg5 C=$C ${vs}cm N=$N rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("$out.json"))
    print("   value %.3e frames/s %.0f upd/frame %.0f e2e %.3e | %s" % (d["value"], d["frames_per_s"], d["updates_per_frame"], d["e2e"]["value"] if d["e2e"] else 0, d["config"]["workload"][:110]))
except Exception as e:
    print("   no line", e); print(open("$out.err").read()[-1500:])
PY
done
