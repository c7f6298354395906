nvidia-smi -L | wc -l
python -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 --no-msa --no-secondary --cpu-pairs 8000 > gpurun_out/bench_r2d_n8.json 2> gpurun_out/bench_r2d_n8.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r2d_n8.err; wc -c gpurun_out/bench_r2d_n8.json
