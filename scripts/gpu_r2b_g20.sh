#!/bin/bash
# round 2, session 2, call 20: sample blocks per pixel chunk (finer work chunks for a GPU's share of a multi-GPU frame): tests, N = 1 unchanged?, one-of-eight emulation
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_g20_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b_g20_pytest.log
O=gpurun_out/r2b_g20.jsonl; : > $O
timeout 600 python scripts/variant_bench.py --config C2 --spp 500 --reps 3 prev:prev default subs4,RTW_CHUNK_SUBS=4 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C2 --spp 100 --reps 4 prev:prev default 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C1 --spp 100 --reps 5 prev:prev default 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C5 --spp 64 --reps 3 prev:prev default 2>&1 | tee -a $O
export RTW_DEBUG_OWN=0,8
timeout 600 python scripts/variant_bench.py --config C2 --spp 500 --reps 5 subs1,RTW_CHUNK_SUBS=1 default subs2,RTW_CHUNK_SUBS=2 subs4,RTW_CHUNK_SUBS=4 subs7,RTW_CHUNK_SUBS=7 2>&1 | tee -a $O
RTW_DEBUG_OWN=0,4 timeout 600 python scripts/variant_bench.py --config C2 --spp 500 --reps 4 subs1,RTW_CHUNK_SUBS=1 default 2>&1 | tee -a $O
timeout 300 python scripts/timeline_probe.py 500 2>&1 | tee gpurun_out/r2b_timeline_own8_subs.jsonl
