#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|furnace|psnr|poisoned|mean radiance|^E  " gpurun_out/pytest_gpu.log | head -30
