#!/bin/bash
# round 2, session 2, call 8: background misses of a GENERATE pass accumulated by one lane; two-launch candidate pre-pass
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_g8_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_g8_pytest.log
O=gpurun_out/r2b_g8.jsonl; : > $O
timeout 600 python scripts/variant_bench.py --config C2 --spp 100 --reps 5 prev:prev default 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C2 --spp 62 --reps 5 prev:prev default 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C2 --spp 8 --reps 6 prev:prev default 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C1 --spp 100 --reps 6 prev:prev default 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C5 --spp 64 --reps 4 prev:prev default 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C2 --spp 500 --reps 2 prev:prev default 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C4 --spp 16 --reps 2 prev:prev default 2>&1 | tee -a $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2b_launches_62spp_g8.csv python scripts/variant_bench.py --child --config C2 --spp 62 --reps 2 --mode wavefront > gpurun_out/r2b_launches_62_g8.log 2>&1
