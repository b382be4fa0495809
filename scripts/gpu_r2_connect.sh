#!/bin/bash
# round 2: CONNECT stage (stackless light-BVH walk, re-batched) vs the light walk inside Lambertian SHADE; C4 / C5 / C2
mkdir -p gpurun_out
O=gpurun_out/r2_connect.jsonl; : > $O
python scripts/variant_bench.py --config C4 --spp 8 --reps 2 connect noconnect,RTW_NO_CONNECT=1 2>&1 | tee -a $O
python scripts/variant_bench.py --config C4 --spp 8 --reps 1 --mode megakernel megakernel 2>&1 | tee -a $O
python scripts/variant_bench.py --config C4 --spp 64 --reps 1 connect noconnect,RTW_NO_CONNECT=1 2>&1 | tee -a $O
python scripts/variant_bench.py --config C5 --spp 64 --reps 3 connect noconnect,RTW_NO_CONNECT=1 2>&1 | tee -a $O
python scripts/variant_bench.py --config C5 --spp 64 --reps 2 --mode megakernel megakernel 2>&1 | tee -a $O
python scripts/variant_bench.py --config C2 --spp 100 --reps 5 default 2>&1 | tee -a $O
python scripts/variant_bench.py --config C2 --spp 100 --reps 2 --mode megakernel megakernel 2>&1 | tee -a $O
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_connect_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_connect_pytest.log
