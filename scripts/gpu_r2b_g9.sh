#!/bin/bash
# round 2, session 2, call 9: one copy of the frame / sincospi / transform code for both halves of the Lambertian mixture sample; two-launch candidate pre-pass
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_g9_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_g9_pytest.log
O=gpurun_out/r2b_g9.jsonl; : > $O
timeout 600 python scripts/variant_bench.py --config C2 --spp 100 --reps 5 prev:prev default nolm:nolm 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C2 --spp 62 --reps 5 prev:prev default nolm:nolm 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C1 --spp 100 --reps 6 prev:prev default nolm:nolm 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C5 --spp 64 --reps 4 prev:prev default nolm:nolm 2>&1 | tee -a $O
timeout 600 python scripts/variant_bench.py --config C2 --spp 100 --reps 3 --mode megakernel prev:prev default nolm:nolm 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C4 --spp 16 --reps 2 prev:prev default 2>&1 | tee -a $O
