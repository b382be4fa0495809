#!/bin/bash
# 2 GPUs: the multi-GPU tests (rtw_render_multi peer / NCCL, rtw_render_rank in two processes) and the bench under torchrun as the driver launches it
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_2gpus.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu_2gpus.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-300 gpurun_out/r2_bench_n2.json; tail -3 gpurun_out/r2_bench_n2.err
