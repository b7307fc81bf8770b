#!/bin/bash
# per-kernel durations of the window path (ncu launch list; cold-cache and serialised - shares, not absolutes)
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-query"
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 200 --csv --log-file gpurun_out/launches_${TAG:-x}.csv $B > gpurun_out/launches_${TAG:-x}.log 2>&1
echo "rc=$?"
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/launches_${TAG:-x}.csv") if l.startswith('"')))
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    agg[r[ki][:60]].append(float(r[vi].replace(",", "")))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-62s n=%3d avg %.1f us" % (k, len(v), sum(v) / len(v) / 1000.0))
PY
