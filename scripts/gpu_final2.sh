#!/bin/bash
# Round-end evidence run (second half of round 1: general scenes, device BVH): everything profiles/ cites that changed.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_n1.json
timeout 900 python scripts/general_configs.py > gpurun_out/general_configs.jsonl 2> gpurun_out/general_configs.err; echo "general rc=$?"
timeout 900 python scripts/lbvh_vs_sah.py 500 64 > gpurun_out/lbvh_vs_sah.log 2>&1; echo "lbvh rc=$?"; tail -7 gpurun_out/lbvh_vs_sah.log
timeout 600 python scripts/stress_configs.py C4 > gpurun_out/stress_c4.jsonl 2>&1; cut -c1-300 gpurun_out/stress_c4.jsonl
# ncu of the general FP32 renderer on a small cornell_box frame (the capture replays the kernel ~40 times)
RTW_GENERAL_SMALL=1 python scripts/general_configs.py cornell_box --once > gpurun_out/once.log 2>&1 &&
RTW_GENERAL_SMALL=1 ncu --set full --clock-control none --import-source on -k regex:render_pool_kernel -c 1 -o gpurun_out/general_cornell_pool -f \
  python scripts/general_configs.py cornell_box --once > gpurun_out/ncu_general.log 2>&1
echo "ncu rc=$?"
