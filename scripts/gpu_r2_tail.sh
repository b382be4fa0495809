#!/bin/bash
# round 2: the end of the path stream — spill at exhaustion + one-path-per-thread resume kernel; 1080p at 62 spp (one of eight GPUs' share), 8 and 500 spp
mkdir -p gpurun_out
O=gpurun_out/r2_tail.jsonl; : > $O
for SPP in 62 8; do
python scripts/variant_bench.py --config C2 --spp $SPP --reps 6 nospill,RTW_SPILL_THRESHOLD=0 spill spill48,RTW_SPILL_THRESHOLD=48 spill24,RTW_SPILL_THRESHOLD=24 nospill_chunk512,RTW_SPILL_THRESHOLD=0,RTW_CHUNK_PATHS=512 2>&1 | tee -a $O
done
python scripts/variant_bench.py --config C2 --spp 500 --reps 3 nospill,RTW_SPILL_THRESHOLD=0 spill 2>&1 | tee -a $O
python scripts/variant_bench.py --config C1 --spp 100 --reps 6 nospill,RTW_SPILL_THRESHOLD=0 spill 2>&1 | tee -a $O
python scripts/variant_bench.py --config C5 --spp 64 --reps 3 nospill,RTW_SPILL_THRESHOLD=0 spill 2>&1 | tee -a $O
python scripts/variant_bench.py --config C4 --spp 16 --reps 2 nospill,RTW_SPILL_THRESHOLD=0 spill 2>&1 | tee -a $O
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tail_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_tail_pytest.log
