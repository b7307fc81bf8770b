python tools/prof_window.py 8 6 2>&1 | tail -n 1
python bench.py --no-cpu-baseline --no-e2e 2>&1 >/dev/null | grep "\[bench\]" | head -1
