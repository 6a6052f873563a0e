# identity experiment: one bulk copy per task into padded (odd 16-byte-unit) slots vs one bulk copy per tile
# (every command under a short timeout: a hung kernel must not eat the box time)
set -x
mkdir -p gpurun_out
rm -f gpurun_out/x_*.json
B9_FORCE_SCATTER=1 timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or handcrafted or fifo or echo" 2>&1 | tail -3
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "scatter parity failed or hung: stop"; exit 0; fi
for rep in 1 2; do
timeout 90 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/x_base_$rep.json 2>> gpurun_out/x.err
B9_FORCE_SCATTER=1 timeout 90 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/x_scatter_$rep.json 2>> gpurun_out/x.err || break
done
B9_FORCE_SCATTER=1 timeout 90 python bench.py --no-cpu-baseline --e2e-steps 2 --adversarial 0 > gpurun_out/x_scatter_clean.json 2>> gpurun_out/x.err
B9_FORCE_SCATTER=1 timeout 90 python bench.py --handler vadd_f32 --tasks 1000000 --no-cpu-baseline --e2e-steps 2 > gpurun_out/x_vadd_scatter.json 2>> gpurun_out/x.err
tail -n 5 gpurun_out/x.err
