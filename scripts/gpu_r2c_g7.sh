#!/bin/bash
# round 2, session 3, call 7: render_background_kernel in small CTAs on a second stream NEXT TO the wavefront kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_g7_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c_g7_pytest.log
O=gpurun_out/r2c_g7.jsonl; : > $O
V="c3:c3 default nosd,RTW_NO_SIDE_STREAM=1"
timeout 900 python scripts/variant_bench.py --config C2 --spp 500 --reps 3 $V 2>&1 | tee -a $O | cut -c1-100
RTW_DEBUG_OWN=0,8 timeout 900 python scripts/variant_bench.py --config C2 --spp 500 --reps 5 $V 2>&1 | tee -a $O | cut -c1-100
timeout 600 python scripts/variant_bench.py --config C1 --spp 100 --reps 5 $V 2>&1 | tee -a $O | cut -c1-100
timeout 600 python scripts/variant_bench.py --config C5 --spp 256 --reps 3 $V 2>&1 | tee -a $O | cut -c1-100
timeout 600 python scripts/variant_bench.py --config C2 --spp 8 --reps 5 $V 2>&1 | tee -a $O | cut -c1-100
