# 2 GPUs: the multi-GPU parity tests, the weak-scaling bench line and the skewed-ingest (NCCL rebalance) line
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --no-cpu-baseline --e2e-steps 5 > gpurun_out/n2.json 2> gpurun_out/n2.err
B9_REBALANCE_TRACE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --no-cpu-baseline --e2e-steps 3 --skew 1.0 > gpurun_out/n2_skew.json 2> gpurun_out/n2_skew.err
grep b9_rebalance gpurun_out/n2_skew.err
