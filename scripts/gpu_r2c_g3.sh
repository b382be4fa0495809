#!/bin/bash
# round 2, session 3, call 3: wavefront instantiated per kind of light list (LightMode) + two-pass light terms in the flat-list kernel + split last chunks
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_g3_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c_g3_pytest.log
O=gpurun_out/r2c_g3.jsonl; : > $O
V="prev:prev default ld:ld"
timeout 900 python scripts/variant_bench.py --config C2 --spp 500 --reps 3 $V 2>&1 | tee -a $O | cut -c1-100
RTW_DEBUG_OWN=0,8 timeout 900 python scripts/variant_bench.py --config C2 --spp 500 --reps 5 $V t4k,RTW_CHEAP_TAIL_PATHS=4096 2>&1 | tee -a $O | cut -c1-100
timeout 600 python scripts/variant_bench.py --config C2 --spp 100 --reps 5 $V 2>&1 | tee -a $O | cut -c1-100
timeout 600 python scripts/variant_bench.py --config C5 --spp 256 --reps 3 $V 2>&1 | tee -a $O | cut -c1-100
timeout 600 python scripts/variant_bench.py --config C1 --spp 100 --reps 5 $V 2>&1 | tee -a $O | cut -c1-100
timeout 600 python scripts/variant_bench.py --config C4 --spp 64 --reps 2 prev:prev default 2>&1 | tee -a $O | cut -c1-100
