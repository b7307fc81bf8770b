python -m pytest tests -q -m gpu 2>&1 | tail -n 3
for w in 8 12 16; do python bench.py --no-cpu-baseline --no-e2e --window $w 2>&1 >/dev/null | grep "\[bench\]" | head -1; done
python tools/prof_window.py 16 4 2>&1 | tail -n 1
