#!/bin/bash
# round 2, session 2, call 10: CTA-wide stage hint (experiment), C4 at HEAD, a full bench.py line
mkdir -p gpurun_out
O=gpurun_out/r2b_g10.jsonl; : > $O
timeout 600 python scripts/variant_bench.py --config C2 --spp 100 --reps 5 default hint32:hint32 hint24:hint24 hint12:hint12 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C5 --spp 64 --reps 4 default hint24:hint24 2>&1 | tee -a $O
timeout 900 python scripts/variant_bench.py --config C4 --spp 16 --reps 2 default 2>&1 | tee -a $O
python bench.py > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2b_bench_n1.json
