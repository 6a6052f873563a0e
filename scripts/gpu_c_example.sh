set -x
timeout 250 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wire.py -m gpu -x -q -k "golden or handcrafted or wire" 2>&1 | tail -5
