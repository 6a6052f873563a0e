# identity: parity tests, then the bench with the in-kernel tail (default), with the second kernel, and an older build
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for rep in 1 2; do
B9GPU_LIB=$PWD/ab_builds/libb9gpu_v0.so timeout 300 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/id_v0_$rep.json 2> gpurun_out/id_v0.err
timeout 300 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/id_tail_$rep.json 2> gpurun_out/id_tail.err
B9_SLOW_KERNEL=1 timeout 300 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/id_2k_$rep.json 2> gpurun_out/id_2k.err
done
B9GPU_LIB=$PWD/ab_builds/libb9gpu_v0.so timeout 300 python bench.py --no-cpu-baseline --e2e-steps 2 --adversarial 0 > gpurun_out/id_v0_clean.json 2> gpurun_out/id_clean.err
timeout 300 python bench.py --no-cpu-baseline --e2e-steps 2 --adversarial 0 > gpurun_out/id_clean.json 2> gpurun_out/id_clean.err
B9_SLOW_KERNEL=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:drain_slow -s 2 -c 1 -o gpurun_out/ncu_slow -f python bench.py --no-cpu-baseline --e2e-steps 1 --steps 2 --warmup 1 > gpurun_out/ncu_slow.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:drain3_kernel -s 2 -c 1 -o gpurun_out/ncu_tail -f python bench.py --no-cpu-baseline --e2e-steps 1 --steps 2 --warmup 1 > gpurun_out/ncu_tail.log 2>&1
tail -n 3 gpurun_out/*.err
