python -m pytest tests/test_sw_gpu.py tests/test_kat_gpu.py -m gpu -x -q 2>&1 | tail -2
python tools/map_probe.py 1000000 2 2>&1 | tail -1
python tools/sw_ragged_probe.py 200000 2 | tail -2
