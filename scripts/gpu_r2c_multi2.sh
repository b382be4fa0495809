#!/bin/bash
# 2 GPUs at the round's last commit: the multi-GPU tests (5 of them are skipped on a 1-GPU box) and bench.py as the driver launches it
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_2gpus.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu_2gpus.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-200 gpurun_out/r2_bench_n2.json
