# what the driver runs at round end, in one call: smoke, the GPU tests, both bench arms
set -x
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
cat gpurun_out/bench_default.json gpurun_out/bench_reference.json | cut -c1-2500
SEL='golden or handcrafted or empty or fifo or cancelled or capacity or corrupted or fast_path or http_body or async or wire'
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wire.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitize_memcheck.log
tail -n 5 gpurun_out/sanitize_memcheck.log
