#!/bin/bash
# Round-end evidence at the round's last commit: tests, smoke, both bench arms, general scenes
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; wc -l gpurun_out/bench_ref.json
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; wc -l gpurun_out/bench_n1.json; cut -c1-200 gpurun_out/bench_n1.json
timeout 900 python scripts/general_configs.py > gpurun_out/general_configs.jsonl 2> gpurun_out/general_configs.err; echo "general rc=$?"
timeout 300 python scripts/general_renderers.py > gpurun_out/general_renderers.log 2>&1; echo "renderers rc=$?"; cat gpurun_out/general_renderers.log
