# seeding experiment: parity of the seeding paths + timing of the mapping pipeline on cfg3 / cfg4 pairs
python -m pytest tests/test_kat_gpu.py tests/test_map_gpu.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5
VG_SEED_DEBUG=2 python tools/map_probe.py 200000 3 2>&1 | tail -9
VG_SEED_DEBUG=1 python tools/map_probe.py 200000 3 4 2>&1 | tail -6
