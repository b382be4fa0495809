#!/bin/bash
# round 2, session 2, call 15: launch-shape sweep, fewer warps with more path slots and registers each
mkdir -p gpurun_out
O=gpurun_out/r2b_g15.jsonl; : > $O
timeout 600 python scripts/variant_bench.py --config C2 --spp 100 --reps 4 s768x96:sweep s640x112:sweep,RTW_WF_SHAPE=3 s576x128:sweep,RTW_WF_SHAPE=6 s512x144:sweep,RTW_WF_SHAPE=7 s640x104:sweep,RTW_WF_SHAPE=8 s448x160:sweep,RTW_WF_SHAPE=9 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C5 --spp 64 --reps 4 s768x96:sweep s640x112:sweep,RTW_WF_SHAPE=3 s576x128:sweep,RTW_WF_SHAPE=6 s512x144:sweep,RTW_WF_SHAPE=7 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C1 --spp 100 --reps 6 s768x96:sweep s640x112:sweep,RTW_WF_SHAPE=3 s576x128:sweep,RTW_WF_SHAPE=6 s512x144:sweep,RTW_WF_SHAPE=7 2>&1 | tee -a $O
