# baseline of a session: GPU tests, then the default bench at N=1
set -x
nvidia-smi -L
python -m pytest tests -m gpu -q -x 2>&1 | tail -15
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2b_n1.json 2> gpurun_out/bench_r2b_n1.err; tail -c 800 gpurun_out/bench_r2b_n1.err
VG_SEED_DEBUG=2 python tools/map_probe.py 200000 3 2>&1 | tail -12
