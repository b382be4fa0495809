#!/bin/bash
# round 2, session 2, call 11: CONNECT livelock fix (tests incl. the new sky-tail test, C4 with a short timeout), compiler-flag variants on C2
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_g11_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_g11_pytest.log
O=gpurun_out/r2b_g11.jsonl; : > $O
timeout 240 python scripts/variant_bench.py --config C4 --spp 16 --reps 2 default 2>&1 | tee -a $O
timeout 400 python scripts/variant_bench.py --config C2 --spp 100 --reps 5 default ptxO2:ptxO2 expensive:expensive O2:O2 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C5 --spp 64 --reps 4 default ptxO2:ptxO2 expensive:expensive 2>&1 | tee -a $O
