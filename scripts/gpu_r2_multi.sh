# Round-2 multi-GPU run (gpurun --gpus N): the NCCL rebalance tests, then the bench line with its rebalance leg.
#   gpurun --gpus 2 -- 'N=2 bash scripts/gpu_r2_multi.sh'
mkdir -p gpurun_out
N=${N:-2}
TAG=${TAG:-m$N}
[ -n "$SKIP_TESTS" ] || timeout 900 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_tests.log
[ -n "$SKIP_DEFAULT" ] || B9_REBALANCE_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
# CONFIGS: the other BASELINE configs to run at this N (default: configs[3]); SKIP_TESTS=1 / SKIP_DEFAULT=1 leave the first two legs out
for c in ${CONFIGS:-3}; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29518 + c)) bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --config $c --sustain-seconds 0.3 > gpurun_out/${TAG}_bench_config$c.json 2> gpurun_out/${TAG}_bench_config$c.err
done
grep "b9_rebalance rank 0" gpurun_out/${TAG}_bench.err 2>/dev/null | tail -12
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench*.json")):
    try:
        d = json.load(open(f)); print(f, "value %.4g e2e %.4g frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"]), "rebalance", json.dumps(d["rebalance"])[:300])
    except Exception as e:
        print(f, "unreadable", e)
PY
