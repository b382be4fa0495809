#!/bin/bash
# 2 GPUs: pixel-ownership partition of one frame (each GPU renders ALL samples of the chunks it owns) against the sample partition
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_2gpus.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu_2gpus.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 (pixels) rc=$?"; cut -c1-200 gpurun_out/r2_bench_n2.json
RTW_MULTI_PARTITION=samples python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-c3 > gpurun_out/r2_bench_n2_samples.json 2> gpurun_out/r2_bench_n2_samples.err; echo "bench n2 (samples) rc=$?"; cut -c1-200 gpurun_out/r2_bench_n2_samples.json
python bench.py --no-cpu-baseline --no-f64 --no-c3 > gpurun_out/r2_bench_n1_check.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_n1_check.json
