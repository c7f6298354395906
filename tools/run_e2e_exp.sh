python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/map_probe.py 1000000 2 2>&1 | tail -1
