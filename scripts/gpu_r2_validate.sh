# Round-2 validation run (one GPU): the whole GPU test suite, A/B bench lines, the default bench line with every leg,
# the reference arm, smoke(), ncu captures. Usage: gpurun -- 'TAG=v1 VARIANTS="cur spec" bash scripts/gpu_r2_validate.sh'
mkdir -p gpurun_out
TAG=${TAG:-v}
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
REPS=2 TAG=${TAG} VARIANTS="${VARIANTS:-cur}" ARGS="--adversarial 0.01 --sustain-seconds 0.2;--adversarial 0 --sustain-seconds 0.2" bash scripts/gpu_r2_ab.sh
timeout 300 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
for h in 2 3 4; do timeout 300 python bench.py --config $h --no-cpu-baseline --sustain-seconds 0.3 > gpurun_out/${TAG}_bench_config$h.json 2> gpurun_out/${TAG}_bench_config$h.err; done
python scripts/bench_lines.py gpurun_out/${TAG}_bench_*.json
if [ -z "$NO_NCU" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:drain3_kernel -s 2 -c 1 -o gpurun_out/${TAG}_ncu_identity -f python bench.py --no-cpu-baseline --e2e-steps 1 --steps 2 --warmup 1 --sustain-seconds 0 > gpurun_out/ncu_a.log 2>&1
timeout 300 python scripts/ncu_traffic.py > gpurun_out/${TAG}_traffic.log 2>&1
# launch list of the same command (per-launch times are cold-cache and serialised: shares of the step, not absolutes)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --no-cpu-baseline --e2e-steps 1 --steps 2 --warmup 1 --sustain-seconds 0 > gpurun_out/ncu_b.log 2>&1
fi
tail -n 3 gpurun_out/${TAG}_*.err | tail -n 30
