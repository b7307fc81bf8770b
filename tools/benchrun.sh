for w in 8 1 4; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --window $w > gpurun_out/bench_w$w.json 2> gpurun_out/bench_w$w.err; done
SAF_K3W_NST=3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_w8n3.json 2> gpurun_out/bench_w8n3.err
tail -n 3 gpurun_out/bench_w8.err
