#!/bin/bash
# round 2, session 1: padded node stride / fmad variants on C2, baselines of C4 / C5 / C1, GPU tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
python scripts/variant_bench.py --config C2 --spp 100 --reps 5 stride64,RTW_SH_NODE_STRIDE=64 stride80 fmad:fmad fmad64:fmad,RTW_SH_NODE_STRIDE=64 2>&1 | tee gpurun_out/r2_s1_variants.jsonl
python scripts/variant_bench.py --config C2 --spp 100 --reps 3 --mode megakernel stride64,RTW_SH_NODE_STRIDE=64 stride80 fmad:fmad 2>&1 | tee -a gpurun_out/r2_s1_variants.jsonl
python scripts/variant_bench.py --config C1 --spp 100 --reps 5 stride64,RTW_SH_NODE_STRIDE=64 stride80 2>&1 | tee -a gpurun_out/r2_s1_variants.jsonl
python scripts/variant_bench.py --config C5 --spp 64 --reps 3 default 2>&1 | tee -a gpurun_out/r2_s1_variants.jsonl
python scripts/variant_bench.py --config C4 --spp 8 --reps 2 default 2>&1 | tee -a gpurun_out/r2_s1_variants.jsonl
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_s1_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_s1_pytest.log
