#!/bin/bash
# round 2, session 3, call 1: the last costly chunks of the queue handed out in eighths (chunk_split_kernel): tests, A/B on the full frame and on one GPU's share of eight
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_g1_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c_g1_pytest.log
O=gpurun_out/r2c_g1.jsonl; : > $O
V="prev:prev default split0,RTW_SPLIT_CHUNKS_PER_WARP=0 split1,RTW_SPLIT_CHUNKS_PER_WARP=1 split4,RTW_SPLIT_CHUNKS_PER_WARP=4 t4k,RTW_CHEAP_TAIL_PATHS=4096 t2k,RTW_CHEAP_TAIL_PATHS=2048 s4t2k,RTW_SPLIT_CHUNKS_PER_WARP=4,RTW_CHEAP_TAIL_PATHS=2048"
timeout 900 python scripts/variant_bench.py --config C2 --spp 500 --reps 3 $V 2>&1 | tee -a $O | cut -c1-120
echo own8
RTW_DEBUG_OWN=0,8 timeout 900 python scripts/variant_bench.py --config C2 --spp 500 --reps 5 $V t1k,RTW_CHEAP_TAIL_PATHS=1024 s4t1k,RTW_SPLIT_CHUNKS_PER_WARP=4,RTW_CHEAP_TAIL_PATHS=1024 2>&1 | tee -a $O | cut -c1-120
echo others
timeout 600 python scripts/variant_bench.py --config C2 --spp 100 --reps 4 prev:prev default 2>&1 | tee -a $O | cut -c1-120
timeout 600 python scripts/variant_bench.py --config C1 --spp 100 --reps 5 prev:prev default 2>&1 | tee -a $O | cut -c1-120
timeout 600 python scripts/variant_bench.py --config C5 --spp 256 --reps 3 prev:prev default 2>&1 | tee -a $O | cut -c1-120
