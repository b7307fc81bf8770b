python -m pytest tests -q -m gpu -k "sequence or window or golden" 2>&1 | tail -n 2
python bench.py --no-cpu-baseline --no-e2e 2>&1 >/dev/null | grep "\[bench\]" | head -1
python tools/prof_window.py 16 4 2>&1 | tail -n 1
