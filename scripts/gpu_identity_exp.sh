# identity: chunked drain (deferred tasks of chunk i beside the main kernel of chunk i+1) vs one main + one second kernel
# (every command under a short timeout: a hung kernel must not eat the box time)
set -x
mkdir -p gpurun_out
rm -f gpurun_out/x_*.json
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "parity failed or hung: stop"; exit 0; fi
for rep in 1 2; do
B9_DRAIN_CHUNKS=1 timeout 90 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/x_c1_$rep.json 2>> gpurun_out/x.err
B9_DRAIN_CHUNKS=2 timeout 90 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/x_c2_$rep.json 2>> gpurun_out/x.err
B9_DRAIN_CHUNKS=4 timeout 90 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/x_c4_$rep.json 2>> gpurun_out/x.err
B9_DRAIN_CHUNKS=7 timeout 90 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/x_c7_$rep.json 2>> gpurun_out/x.err
done
B9_DRAIN_CHUNKS=4 timeout 90 python bench.py --no-cpu-baseline --e2e-steps 2 --adversarial 0 > gpurun_out/x_c4_clean.json 2>> gpurun_out/x.err
B9_DRAIN_CHUNKS=4 timeout 90 python bench.py --handler json_sum --tasks 300000 --no-cpu-baseline --e2e-steps 2 > gpurun_out/x_json.json 2>> gpurun_out/x.err
tail -n 5 gpurun_out/x.err
