# identity bench (current build vs ab_builds/libb9gpu_v0.so if present), then compute-sanitizer over the small parity tests
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
  [ -f ab_builds/libb9gpu_v0.so ] && B9GPU_LIB=$PWD/ab_builds/libb9gpu_v0.so timeout 300 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/id_v0_$rep.json 2>> gpurun_out/bench.err
  timeout 300 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/id_new_$rep.json 2>> gpurun_out/bench.err
done
SEL='golden or handcrafted or empty or fifo or cancelled or capacity or corrupted or fast_path'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wire.py -m gpu -x -q -k "$SEL or wire" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitize_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or handcrafted" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitize_racecheck.log
tail -n 6 gpurun_out/sanitize_memcheck.log gpurun_out/sanitize_racecheck.log
