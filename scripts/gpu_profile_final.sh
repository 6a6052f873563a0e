# Round profile run (one GPU): launch list of the bench command, full captures of every drain kernel.
# Reports land in gpurun_out/; scripts/ncu_summary.py condenses them for profiles/.
set -x
mkdir -p gpurun_out
R=${ROUND:-r1}
# 1. launch list of the default bench command (durations are cold-cache and serialised: shares, not absolutes)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${R}_launches_bench.log 2>&1
# 2. identity: main + second kernel, one launch each, full set
timeout 300 ncu --set full --clock-control none --import-source on -k regex:drain3_kernel -s 2 -c 1 -o gpurun_out/${R}_ncu_identity_main -f python bench.py --no-cpu-baseline --e2e-steps 1 --steps 2 --warmup 1 > gpurun_out/ncu_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:drain_slow -s 2 -c 1 -o gpurun_out/${R}_ncu_identity_slow -f python bench.py --no-cpu-baseline --e2e-steps 1 --steps 2 --warmup 1 > gpurun_out/ncu_b.log 2>&1
# 3. the other handlers' drain kernels
for h in crc32:500000 vadd_f32:1000000 json_sum:300000; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:drain3_kernel -s 2 -c 1 -o gpurun_out/${R}_ncu_${h%%:*} -f python bench.py --handler ${h%%:*} --tasks ${h##*:} --no-cpu-baseline --e2e-steps 1 --steps 2 --warmup 1 > gpurun_out/ncu_c.log 2>&1
done
ls -la gpurun_out | tail -n 12
