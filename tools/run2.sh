nvidia-smi -L
python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-msa --no-secondary --cpu-pairs 8000 > gpurun_out/bench_r2c_n2.json 2> gpurun_out/bench_r2c_n2.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_r2c_n2.err; wc -c gpurun_out/bench_r2c_n2.json
