python -m pytest tests -q -m gpu -k "slab or sequence" 2>&1 | tail -n 3
export SAF_DEBUG_REACH=1
python tools/prof_rooms.py 2 0 48 2>&1 | tail -n 2
python tools/prof_rooms.py 4 1 48 2>&1 | tail -n 2
python tools/prof_rooms.py 8 3 48 2>&1 | tail -n 2
