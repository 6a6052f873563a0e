set -x
timeout 200 python -m pytest tests/test_gpu_sdk.py tests/test_gpu_c_example.py -m gpu -x -q 2>&1 | tail -4
