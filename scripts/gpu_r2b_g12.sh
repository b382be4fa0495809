#!/bin/bash
# round 2, session 2, call 12: background-only chunks beyond a tail leave the wavefront's queue for render_background_kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_g12_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_g12_pytest.log
O=gpurun_out/r2b_g12.jsonl; : > $O
V="off,RTW_CHEAP_TAIL_PATHS=-1 default t0,RTW_CHEAP_TAIL_PATHS=0 t2k,RTW_CHEAP_TAIL_PATHS=2048 t32k,RTW_CHEAP_TAIL_PATHS=32768"
timeout 400 python scripts/variant_bench.py --config C2 --spp 100 --reps 5 $V 2>&1 | tee -a $O
timeout 400 python scripts/variant_bench.py --config C2 --spp 62 --reps 5 $V 2>&1 | tee -a $O
timeout 400 python scripts/variant_bench.py --config C2 --spp 8 --reps 6 $V 2>&1 | tee -a $O
timeout 400 python scripts/variant_bench.py --config C1 --spp 100 --reps 6 $V 2>&1 | tee -a $O
timeout 400 python scripts/variant_bench.py --config C5 --spp 64 --reps 4 off,RTW_CHEAP_TAIL_PATHS=-1 default 2>&1 | tee -a $O
timeout 400 python scripts/variant_bench.py --config C2 --spp 500 --reps 2 off,RTW_CHEAP_TAIL_PATHS=-1 default 2>&1 | tee -a $O
timeout 240 python scripts/variant_bench.py --config C4 --spp 16 --reps 2 off,RTW_CHEAP_TAIL_PATHS=-1 default 2>&1 | tee -a $O
