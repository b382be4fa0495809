#!/bin/bash
# round 2, session 2, call 7: work chunks = 32 pixels x a block of samples (a lane stays on one pixel): GPU tests, A/B against the previous library
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_g7_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_g7_pytest.log
O=gpurun_out/r2b_g7.jsonl; : > $O
timeout 600 python scripts/variant_bench.py --config C2 --spp 100 --reps 5 prev:prev default s4,RTW_CHUNK_SAMPLES=4 s8,RTW_CHUNK_SAMPLES=8 s32,RTW_CHUNK_SAMPLES=32 s64,RTW_CHUNK_SAMPLES=64 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C2 --spp 62 --reps 5 prev:prev default s4,RTW_CHUNK_SAMPLES=4 s8,RTW_CHUNK_SAMPLES=8 s32,RTW_CHUNK_SAMPLES=32 s62,RTW_CHUNK_SAMPLES=62 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C2 --spp 8 --reps 6 prev:prev default s4,RTW_CHUNK_SAMPLES=4 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C1 --spp 100 --reps 6 prev:prev default s8,RTW_CHUNK_SAMPLES=8 s16,RTW_CHUNK_SAMPLES=16 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C5 --spp 64 --reps 4 prev:prev default 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C2 --spp 500 --reps 2 prev:prev default 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C4 --spp 16 --reps 2 prev:prev default 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config cornell_box --spp 256 --reps 2 prev:prev default 2>&1 | tee -a $O
