#!/bin/bash
# round 2, session 2, call 6: which of the code-size changes cost time (one switch each), against the previous commit's library
mkdir -p gpurun_out
O=gpurun_out/r2b_g6.jsonl; : > $O
timeout 600 python scripts/variant_bench.py --config C2 --spp 100 --reps 5 prev:prev default nonbr,RTW_NO_SELF_HIT_NEIGHBOURS=1 sf:sf cr:cr po:po sf_cr:sf_cr all:all 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C5 --spp 64 --reps 4 prev:prev default sf:sf cr:cr po:po 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C1 --spp 100 --reps 6 prev:prev default sf:sf cr:cr po:po 2>&1 | tee -a $O
