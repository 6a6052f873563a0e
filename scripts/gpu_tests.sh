# the GPU test suite, as the driver runs it at round end
set -x
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
