for w in 16 24; do
SAF_K3W_WARPS=$w python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_w8_$w.json 2> gpurun_out/bench_w8_$w.err
grep "\[bench\]" gpurun_out/bench_w8_$w.err
done
python tools/prof_window.py 8 6 2>&1 | tail -n 1
