# quick GPU check: parity tests + bench lines (no profiler); short timeouts on everything
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 120 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/h_identity.json 2> gpurun_out/h_identity.err
timeout 120 python bench.py --no-cpu-baseline --e2e-steps 2 --cancelled 0.05 > gpurun_out/h_cancel5.json 2> gpurun_out/h_cancel5.err
timeout 120 python bench.py --no-cpu-baseline --e2e-steps 2 --cancelled 0.000002 > gpurun_out/h_cancel1.json 2> gpurun_out/h_cancel1.err
timeout 120 python bench.py --handler json_sum --tasks 300000 --no-cpu-baseline --e2e-steps 2 --cancelled 0.05 > gpurun_out/h_json_cancel5.json 2> gpurun_out/h_json_cancel5.err
tail -n 3 gpurun_out/h_*.err
