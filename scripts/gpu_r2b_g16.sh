#!/bin/bash
# round 2, session 2, call 16: stage selection biased towards full EXTEND batches; size of the background tail kept in the queue
mkdir -p gpurun_out
O=gpurun_out/r2b_g16.jsonl; : > $O
timeout 600 python scripts/variant_bench.py --config C2 --spp 100 --reps 4 default extb1:extb1 extb2:extb2 extb3:extb3 t4k,RTW_CHEAP_TAIL_PATHS=4096 t16k,RTW_CHEAP_TAIL_PATHS=16384 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C5 --spp 64 --reps 4 default extb1:extb1 extb2:extb2 extb3:extb3 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C2 --spp 62 --reps 4 default t4k,RTW_CHEAP_TAIL_PATHS=4096 t16k,RTW_CHEAP_TAIL_PATHS=16384 2>&1 | tee -a $O
