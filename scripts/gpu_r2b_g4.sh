#!/bin/bash
# round 2, session 2, call 4: neighbour lists for the own-sphere shortcut (tests + A/B), per-warp timeline of the tail with / without the queue order
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_g4_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_g4_pytest.log
O=gpurun_out/r2b_g4.jsonl; : > $O
timeout 300 python scripts/variant_bench.py --config C2 --spp 100 --reps 5 nonbr,RTW_NO_SELF_HIT_NEIGHBOURS=1 default 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C5 --spp 64 --reps 4 nonbr,RTW_NO_SELF_HIT_NEIGHBOURS=1 default 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C1 --spp 100 --reps 6 nonbr,RTW_NO_SELF_HIT_NEIGHBOURS=1 default 2>&1 | tee -a $O
T=gpurun_out/r2b_timeline.jsonl; : > $T
echo '{"variant": "default"}' >> $T; timeout 300 python scripts/timeline_probe.py 62 8 500 2>&1 | tee -a $T
echo '{"variant": "RTW_NO_CHUNK_ORDER=1"}' >> $T; RTW_NO_CHUNK_ORDER=1 timeout 300 python scripts/timeline_probe.py 62 8 2>&1 | tee -a $T
