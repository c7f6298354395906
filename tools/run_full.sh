python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python tools/sw_ragged_probe.py 200000 2 | tail -2
