#!/bin/bash
# round 2, session 2, call 19: what one of eight GPUs does (RTW_DEBUG_OWN=0,8 on one GPU): tail size, timeline, launch list
mkdir -p gpurun_out
export RTW_DEBUG_OWN=0,8
O=gpurun_out/r2b_g19.jsonl; : > $O
timeout 600 python scripts/variant_bench.py --config C2 --spp 500 --reps 5 default t0,RTW_CHEAP_TAIL_PATHS=0 t1k,RTW_CHEAP_TAIL_PATHS=1024 t2k,RTW_CHEAP_TAIL_PATHS=2048 t4k,RTW_CHEAP_TAIL_PATHS=4096 t16k,RTW_CHEAP_TAIL_PATHS=16384 own3,RTW_DEBUG_OWN=3,8 own7,RTW_DEBUG_OWN=7,8 2>&1 | tee -a $O
T=gpurun_out/r2b_timeline_own8.jsonl; : > $T
timeout 300 python scripts/timeline_probe.py 500 2>&1 | tee -a $T
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2b_launches_own8.csv python scripts/variant_bench.py --child --config C2 --spp 500 --reps 2 --mode wavefront > gpurun_out/r2b_launches_own8.log 2>&1
