#!/bin/bash
# round 2: per-pixel candidate lists for the camera rays — A/B (RTW_NO_PRIMARY_CANDIDATES=1 = walk the tree), images must be identical
mkdir -p gpurun_out
O=gpurun_out/r2_candidates.jsonl; : > $O
python scripts/variant_bench.py --config C2 --spp 100 --reps 5 candidates tree,RTW_NO_PRIMARY_CANDIDATES=1 2>&1 | tee -a $O
python scripts/variant_bench.py --config C2 --spp 100 --reps 2 --mode megakernel candidates tree,RTW_NO_PRIMARY_CANDIDATES=1 2>&1 | tee -a $O
python scripts/variant_bench.py --config C1 --spp 100 --reps 5 candidates tree,RTW_NO_PRIMARY_CANDIDATES=1 2>&1 | tee -a $O
python scripts/variant_bench.py --config C5 --spp 64 --reps 3 candidates tree,RTW_NO_PRIMARY_CANDIDATES=1 2>&1 | tee -a $O
python scripts/variant_bench.py --config C4 --spp 16 --reps 2 candidates tree,RTW_NO_PRIMARY_CANDIDATES=1 2>&1 | tee -a $O
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_cand_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_cand_pytest.log
