#!/bin/bash
# round 2, session 3, call 2: formulations of the piece decode in the wavefront's queue fetch (the kernel moves by +-0.6 % with 8 instructions), deferred light terms
mkdir -p gpurun_out
O=gpurun_out/r2c_g2.jsonl; : > $O
V="prev:prev default pf1:pf1 pf2:pf2 pf3:pf3 ld:ld ld3:ld3"
timeout 900 python scripts/variant_bench.py --config C2 --spp 500 --reps 3 $V 2>&1 | tee -a $O | cut -c1-100
echo own8
RTW_DEBUG_OWN=0,8 timeout 900 python scripts/variant_bench.py --config C2 --spp 500 --reps 5 $V 2>&1 | tee -a $O | cut -c1-100
echo others
timeout 600 python scripts/variant_bench.py --config C2 --spp 100 --reps 5 $V 2>&1 | tee -a $O | cut -c1-100
timeout 600 python scripts/variant_bench.py --config C5 --spp 256 --reps 3 prev:prev default pf3:pf3 ld:ld 2>&1 | tee -a $O | cut -c1-100
