python -m pytest tests/test_parity_gpu.py -q -k "sequence or golden or depth_aware or slabs or cfg1" 2>&1 | tail -n 5
python tools/prof_window.py 8 6 > gpurun_out/prof_w8.txt 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --window 8 > gpurun_out/bench_w8.json 2> gpurun_out/bench_w8.err
ncu --set full --import-source on --clock-control none -k regex:feature_accumulate_window -s 3 -c 2 -o gpurun_out/r01_k3w -f python tools/prof_window.py 8 4 > gpurun_out/ncu_k3w.log 2>&1
tail -n 2 gpurun_out/prof_w8.txt; cat gpurun_out/bench_w8.json | cut -c1-200
