#!/bin/bash
# round 2, session 2, call 2: own-sphere shortcut (closest_prim_self) and costly-chunks-first queue order: GPU tests, A/B on C2 / C1 / C5, fixed cost
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_g2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_g2_pytest.log
O=gpurun_out/r2b_g2.jsonl; : > $O
timeout 300 python scripts/variant_bench.py --config C2 --spp 100 --reps 5 both_off,RTW_NO_SELF_HIT=1,RTW_NO_CHUNK_ORDER=1 default noself,RTW_NO_SELF_HIT=1 noorder,RTW_NO_CHUNK_ORDER=1 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C2 --spp 62 --reps 6 both_off,RTW_NO_SELF_HIT=1,RTW_NO_CHUNK_ORDER=1 default noself,RTW_NO_SELF_HIT=1 noorder,RTW_NO_CHUNK_ORDER=1 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C2 --spp 8 --reps 6 both_off,RTW_NO_SELF_HIT=1,RTW_NO_CHUNK_ORDER=1 default noself,RTW_NO_SELF_HIT=1 noorder,RTW_NO_CHUNK_ORDER=1 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C1 --spp 100 --reps 6 both_off,RTW_NO_SELF_HIT=1,RTW_NO_CHUNK_ORDER=1 default noself,RTW_NO_SELF_HIT=1 noorder,RTW_NO_CHUNK_ORDER=1 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C5 --spp 64 --reps 3 both_off,RTW_NO_SELF_HIT=1,RTW_NO_CHUNK_ORDER=1 default noself,RTW_NO_SELF_HIT=1 noorder,RTW_NO_CHUNK_ORDER=1 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C2 --spp 500 --reps 2 both_off,RTW_NO_SELF_HIT=1,RTW_NO_CHUNK_ORDER=1 default 2>&1 | tee -a $O
timeout 300 python scripts/fixed_cost_probe.py 2>&1 | tee gpurun_out/r2b_fixed_cost_probe.jsonl
