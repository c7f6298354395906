set -x
nvidia-smi -L
python -m pytest tests -m gpu -q 2>&1 | tail -40
python bench.py --steps 3 --warmup 3 --pairs 200000 --secondary-pairs 50000 --cpu-pairs 16000 > gpurun_out/bench_r2a_n1.json 2> gpurun_out/bench_r2a_n1.err; tail -c 600 gpurun_out/bench_r2a_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --pairs 200000 --no-msa --no-secondary --cpu-pairs 8000 > gpurun_out/bench_r2a_n2.json 2> gpurun_out/bench_r2a_n2.err; tail -c 1500 gpurun_out/bench_r2a_n2.err
