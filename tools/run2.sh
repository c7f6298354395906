nvidia-smi -L | wc -l
python -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -5
