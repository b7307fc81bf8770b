python bench.py --workload cfg3 --no-cpu-baseline > gpurun_out/bench_r01_cfg3.json 2> gpurun_out/bench_r01_cfg3.err; grep "\[bench\]" gpurun_out/bench_r01_cfg3.err; tail -n 1 gpurun_out/bench_r01_cfg3.err | cut -c1-200
python -c "
import json
d=json.load(open('gpurun_out/bench_r01_cfg3.json')); print(d['value'], d['frames_per_s'], d['updates_per_frame'], d['timed_region_attempts_ms'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['whole_step_frac'])
"
