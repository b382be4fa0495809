#!/bin/bash
# round 2: CTA-level hand-over of the last paths (donor warps -> collector warps) at the end of the stream
mkdir -p gpurun_out
O=gpurun_out/r2_drain.jsonl; : > $O
for SPP in 62 8; do
timeout 300 python scripts/variant_bench.py --config C2 --spp $SPP --reps 6 off,RTW_DRAIN_COLLECTORS=0 c4m24 c4m48,RTW_DRAIN_MAX=48 c2m24,RTW_DRAIN_COLLECTORS=2 c8m24,RTW_DRAIN_COLLECTORS=8 c4m12,RTW_DRAIN_MAX=12 c4m96,RTW_DRAIN_MAX=96 2>&1 | tee -a $O
done
timeout 300 python scripts/variant_bench.py --config C2 --spp 250 --reps 3 off,RTW_DRAIN_COLLECTORS=0 c4m24 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C1 --spp 100 --reps 6 off,RTW_DRAIN_COLLECTORS=0 c4m24 c4m48,RTW_DRAIN_MAX=48 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C5 --spp 64 --reps 3 off,RTW_DRAIN_COLLECTORS=0 c4m24 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C4 --spp 16 --reps 2 off,RTW_DRAIN_COLLECTORS=0 c4m24 2>&1 | tee -a $O
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_drain_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_drain_pytest.log
