# window kernel on cfg4: shared-memory capacity for the surviving LongHits (capRs) vs resident warps
set -x
for w in 32 28 24 20; do VG_SEED_WINWARPS=$w VG_SEED_DEBUG=2 python tools/map_probe.py 200000 2 4 2>&1 | tail -3; done
VG_SEED_WINWARPS=24 VG_SEED_DEBUG=1 python tools/map_probe.py 200000 2 3 2>&1 | tail -2
python tools/sw_ragged_probe.py 200000 2 | tail -2
