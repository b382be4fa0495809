#!/bin/bash
# round 2, session 2, call 13: three ways of telling the wavefront where its queue ends (register / load per fetch / flag in the order entry)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_g13_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_g13_pytest.log
O=gpurun_out/r2b_g13.jsonl; : > $O
V="prev:prev default qe0:qe0 qe1:qe1 off,RTW_CHEAP_TAIL_PATHS=-1"
timeout 400 python scripts/variant_bench.py --config C2 --spp 100 --reps 5 $V 2>&1 | tee -a $O
timeout 400 python scripts/variant_bench.py --config C2 --spp 62 --reps 5 $V 2>&1 | tee -a $O
timeout 400 python scripts/variant_bench.py --config C1 --spp 100 --reps 6 $V 2>&1 | tee -a $O
timeout 400 python scripts/variant_bench.py --config C5 --spp 64 --reps 4 $V 2>&1 | tee -a $O
timeout 400 python scripts/variant_bench.py --config C2 --spp 500 --reps 2 prev:prev default qe1:qe1 2>&1 | tee -a $O
