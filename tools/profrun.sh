export SAF_BENCH_STEPS=1
python bench.py > gpurun_out/bench_r01_h.json 2> gpurun_out/bench_r01_h.err; grep "\[bench\]" gpurun_out/bench_r01_h.err; tail -n 1 gpurun_out/bench_r01_h.err | cut -c1-200
python -c "
import json
d=json.load(open('gpurun_out/bench_r01_h.json')); print(d['value'], d['frames_per_s'], d['timed_region_attempts_ms'], d['e2e']['value'], d['e2e']['frames_per_s'], d['roofline']['frac'], d['cpu_baseline']['value'])
"
