#!/bin/bash
# round 2, session 2, call 14: launch-shape sweep of the wavefront kernel at the round's code (threads per CTA x path slots per warp)
mkdir -p gpurun_out
O=gpurun_out/r2b_g14.jsonl; : > $O
timeout 600 python scripts/variant_bench.py --config C2 --spp 100 --reps 4 default s768x96:sweep s896x64:sweep,RTW_WF_SHAPE=1 s832x80:sweep,RTW_WF_SHAPE=2 s640x112:sweep,RTW_WF_SHAPE=3 s704x96:sweep,RTW_WF_SHAPE=4 s768x88:sweep,RTW_WF_SHAPE=5 2>&1 | tee -a $O
