set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --handler crc32 --tasks 200000 --no-cpu-baseline --e2e-steps 2 > gpurun_out/h_crc32.json 2> gpurun_out/h_crc32.err
timeout 300 python bench.py --handler vadd_f32 --tasks 200000 --no-cpu-baseline --e2e-steps 2 > gpurun_out/h_vadd.json 2> gpurun_out/h_vadd.err
timeout 300 python bench.py --handler json_sum --tasks 100000 --no-cpu-baseline --e2e-steps 2 > gpurun_out/h_json.json 2> gpurun_out/h_json.err
cat gpurun_out/h_*.json
for h in crc32 vadd_f32 json_sum; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:drain3_kernel -s 2 -c 1 -o gpurun_out/ncu_$h -f python bench.py --handler $h --tasks 100000 --no-cpu-baseline --e2e-steps 1 --steps 2 --warmup 1 > gpurun_out/ncu_$h.log 2>&1
done
ls -la gpurun_out
