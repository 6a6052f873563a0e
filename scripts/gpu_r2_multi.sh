# Round-2 multi-GPU run (gpurun --gpus N): the NCCL rebalance tests, then the bench line with its rebalance leg.
#   gpurun --gpus 2 -- 'N=2 bash scripts/gpu_r2_multi.sh'
mkdir -p gpurun_out
N=${N:-2}
TAG=${TAG:-m$N}
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_tests.log
B9_REBALANCE_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --config 3 --sustain-seconds 0.3 > gpurun_out/${TAG}_bench_config3.json 2> gpurun_out/${TAG}_bench_config3.err
grep "b9_rebalance rank 0" gpurun_out/${TAG}_bench.err | tail -12
tail -4 gpurun_out/${TAG}_bench.err gpurun_out/${TAG}_bench_config3.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench.json", "gpurun_out/${TAG}_bench_config3.json"):
    try:
        d = json.load(open(f)); print(f, "value %.4g e2e %.4g" % (d["value"], d["e2e"]["value"]), "rebalance", json.dumps(d["rebalance"])[:400])
    except Exception as e:
        print(f, "unreadable", e)
PY
