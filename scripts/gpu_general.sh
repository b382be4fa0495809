#!/bin/bash
# general-scene parity, the whole GPU suite, general-scene throughput, ncu capture of the general kernel on cornell_box
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python scripts/general_configs.py > gpurun_out/general_configs.jsonl 2> gpurun_out/general_configs.err; echo "configs rc=$?"
cat gpurun_out/general_configs.jsonl
timeout 300 python scripts/general_configs.py cornell_box --once > gpurun_out/once.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_mega_kernel -c 1 -o gpurun_out/general_cornell \
  python scripts/general_configs.py cornell_box --once > gpurun_out/ncu_general.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
