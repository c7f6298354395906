python -m pytest tests/test_sw_gpu.py -m gpu -x -q 2>&1 | tail -2
python tools/map_probe.py 1000000 3 2>&1 | tail -2
VG_SW_NOBIN=1 python tools/map_probe.py 1000000 3 2>&1 | tail -2
