# compute-sanitizer over the small GPU tests (round 2: tile copy, escape check, kernel tail, pickle path, sink, host paths)
mkdir -p gpurun_out
SEL='golden or handcrafted or empty or fifo or cancelled or capacity or corrupted or fast_path or escaped'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wire.py tests/test_gpu_function.py tests/test_gpu_sink.py -m gpu -x -q -k "($SEL or wire or function or foreign or mirror or fetch_object) and not at_size and not full_size and not per_gpu" > gpurun_out/r2_sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_sanitize_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_function.py -m gpu -x -q -k "golden or handcrafted or foreign" > gpurun_out/r2_sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2_sanitize_racecheck.log
tail -n 8 gpurun_out/r2_sanitize_memcheck.log gpurun_out/r2_sanitize_racecheck.log
