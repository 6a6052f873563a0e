# identity: hybrid static/ticket tile assignment against pure work stealing, on one box; all parity tests first
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
B9_STATIC_ROUNDS=0 timeout 300 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/x_dyn_$rep.json 2>> gpurun_out/x.err
timeout 300 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/x_hybrid_$rep.json 2>> gpurun_out/x.err
done
B9_STATIC_ROUNDS=0 timeout 300 python bench.py --handler vadd_f32 --tasks 1000000 --no-cpu-baseline --e2e-steps 2 > gpurun_out/x_vadd_dyn.json 2>> gpurun_out/x.err
timeout 300 python bench.py --handler vadd_f32 --tasks 1000000 --no-cpu-baseline --e2e-steps 2 > gpurun_out/x_vadd_hybrid.json 2>> gpurun_out/x.err
timeout 300 python bench.py --no-cpu-baseline --e2e-steps 2 --adversarial 0 > gpurun_out/x_clean.json 2>> gpurun_out/x.err
tail -n 5 gpurun_out/x.err
