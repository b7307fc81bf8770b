python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 2 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_f.json 2> gpurun_out/bench_r01_f.err; tail -n 2 gpurun_out/bench_r01_f.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_ref.json 2> gpurun_out/bench_r01_ref.err; cut -c1-300 gpurun_out/bench_r01_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_window.csv python bench.py --steps 2 --warmup 1 --frames-per-step 48 --no-cpu-baseline --no-e2e --pool 96 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:tsdf_update -s 3 -c 2 -o gpurun_out/r01_k2w -f python tools/prof_window.py 8 4 > gpurun_out/ncu_k2w.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:feature_accumulate_window -s 3 -c 2 -o gpurun_out/r01_k3w -f python tools/prof_window.py 8 4 > gpurun_out/ncu_k3w.log 2>&1
python -c "
import json; d=json.load(open('gpurun_out/bench_r01_f.json')); print(d['value'], d['e2e']['value'], d['e2e']['frames_per_s'], d['roofline']['frac'], d['cpu_baseline']['value'])"
