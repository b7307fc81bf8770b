python -m pytest tests/test_parity_gpu.py -q -k "sequence" 2>&1 | tail -n 3
python tools/prof_window.py 8 6 2>&1 | tail -n 1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_p1.json 2> gpurun_out/bench_p1.err; grep "\[bench\]" gpurun_out/bench_p1.err
