python bench.py --steps 10 --warmup 3 > gpurun_out/b1.json 2> gpurun_out/b1.err
python bench.py --steps 10 --warmup 3 --window 1 --no-cpu-baseline > gpurun_out/b2.json 2> gpurun_out/b2.err
for f in b1 b2; do tail -n 2 gpurun_out/$f.err; cut -c1-150 gpurun_out/$f.json; done
