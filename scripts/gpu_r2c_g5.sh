#!/bin/bash
# round 2, session 3, call 5: code-size combinations (Philox out of line, light loop unroll, merged Lambertian halves)
mkdir -p gpurun_out
O=gpurun_out/r2c_g5.jsonl; : > $O
V="default po:po po_u1:po_u1 po_u2:po_u2 po_lm:po_lm po_u1_lm:po_u1_lm po_u2_lm:po_u2_lm lm:lm u1:u1"
timeout 900 python scripts/variant_bench.py --config C2 --spp 500 --reps 3 $V 2>&1 | tee -a $O | cut -c1-100
timeout 600 python scripts/variant_bench.py --config C5 --spp 256 --reps 3 default po:po po_lm:po_lm po_u1_lm:po_u1_lm 2>&1 | tee -a $O | cut -c1-100
timeout 600 python scripts/variant_bench.py --config C1 --spp 100 --reps 5 default po:po po_lm:po_lm po_u1_lm:po_u1_lm 2>&1 | tee -a $O | cut -c1-100
timeout 600 python scripts/variant_bench.py --config C4 --spp 64 --reps 2 default po:po po_lm:po_lm 2>&1 | tee -a $O | cut -c1-100
timeout 600 python scripts/variant_bench.py --config cornell_box --spp 256 --reps 2 default po:po po_lm:po_lm 2>&1 | tee -a $O | cut -c1-100
