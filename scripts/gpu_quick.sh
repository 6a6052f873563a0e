# quick GPU check: parity tests + one bench line per handler (no profiler); short timeouts on everything
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 120 python bench.py --handler crc32 --tasks 500000 --no-cpu-baseline --e2e-steps 2 > gpurun_out/h_crc32.json 2> gpurun_out/h_crc32.err
timeout 120 python bench.py --handler vadd_f32 --tasks 1000000 --no-cpu-baseline --e2e-steps 2 > gpurun_out/h_vadd.json 2> gpurun_out/h_vadd.err
timeout 120 python bench.py --handler json_sum --tasks 300000 --no-cpu-baseline --e2e-steps 2 > gpurun_out/h_json.json 2> gpurun_out/h_json.err
timeout 120 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/h_identity.json 2> gpurun_out/h_identity.err
timeout 120 python bench.py --no-cpu-baseline --e2e-steps 2 --cancelled 0.05 > gpurun_out/h_cancel5.json 2> gpurun_out/h_cancel5.err
B9_STATIC_ROUNDS=0 timeout 120 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/h_identity_dyn.json 2> gpurun_out/h_identity_dyn.err
tail -n 3 gpurun_out/h_*.err
