# 8 GPUs: the weak-scaling bench line the driver asks for (identity, 1M tasks per GPU)
set -x
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --no-cpu-baseline --e2e-steps 5 > gpurun_out/n8.json 2> gpurun_out/n8.err
tail -n 3 gpurun_out/n8.err
