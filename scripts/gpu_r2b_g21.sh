#!/bin/bash
# C3 (4K, 1024 spp) twice through the stress script and once through bench.py --workload C3: the two paths should agree
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv,noheader
timeout 600 python scripts/stress_configs.py C3 > gpurun_out/r2b_c3_stress_a.jsonl 2>&1; cat gpurun_out/r2b_c3_stress_a.jsonl | cut -c1-300
nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv,noheader
timeout 600 python bench.py --workload C3 --steps 2 --warmup 1 --no-cpu-baseline --no-f64 > gpurun_out/r2b_c3_bench.json 2>/dev/null; cut -c1-400 gpurun_out/r2b_c3_bench.json
timeout 600 python scripts/stress_configs.py C3 > gpurun_out/r2b_c3_stress_b.jsonl 2>&1; cat gpurun_out/r2b_c3_stress_b.jsonl | cut -c1-300
