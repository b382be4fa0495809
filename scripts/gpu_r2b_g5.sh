#!/bin/bash
# round 2, session 2, call 5: hot code back under the instruction cache (out-of-line Philox, one accumulate site, rolled candidate / neighbour loops)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_g5_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_g5_pytest.log
O=gpurun_out/r2b_g5.jsonl; : > $O
timeout 400 python scripts/variant_bench.py --config C2 --spp 100 --reps 5 prev:prev default nonbr,RTW_NO_SELF_HIT_NEIGHBOURS=1 philox_inline:philox_inline philox_inline_nonbr:philox_inline,RTW_NO_SELF_HIT_NEIGHBOURS=1 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C5 --spp 64 --reps 4 prev:prev default nonbr,RTW_NO_SELF_HIT_NEIGHBOURS=1 philox_inline:philox_inline 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C1 --spp 100 --reps 6 prev:prev default nonbr,RTW_NO_SELF_HIT_NEIGHBOURS=1 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C2 --spp 62 --reps 5 prev:prev default 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C4 --spp 16 --reps 2 prev:prev default 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config cornell_box --spp 256 --reps 2 prev:prev default 2>&1 | tee -a $O
