#!/bin/bash
# general-scene parity + the whole GPU suite + a short headline bench (regression check of the sphere path)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_general.py -m gpu -q -x > gpurun_out/pytest_general.log 2>&1; echo "general rc=$?"
tail -30 gpurun_out/pytest_general.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "parity rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
cat gpurun_out/bench_quick.json | head -c 600
