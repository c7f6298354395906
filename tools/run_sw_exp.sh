for r in 4 6 8; do VG_SW_MINR=$r python tools/sw_ragged_probe.py 200000 2 | tail -1; done
python tools/sw_ragged_probe.py 200000 2
python -m pytest tests/test_sw_gpu.py tests/test_map_gpu.py -m gpu -x -q 2>&1 | tail -3
