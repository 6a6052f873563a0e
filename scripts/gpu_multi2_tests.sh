# 2 GPUs: only the multi-GPU parity tests (skewed push, cancelled tasks travelling, rebalance, drain)
set -x
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
