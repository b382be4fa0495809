#!/bin/bash
# round 2, session 3, call 4: SFU reciprocal for the fast path's three IEEE divisions, Philox out of line, unroll factor of the light loop's first pass; ncu of the new default
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_g4_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c_g4_pytest.log
O=gpurun_out/r2c_g4.jsonl; : > $O
V="c1:c1 default ieee:ieee po:po u1:u1 u2:u2 u8:u8"
timeout 900 python scripts/variant_bench.py --config C2 --spp 500 --reps 3 $V 2>&1 | tee -a $O | cut -c1-100
RTW_DEBUG_OWN=0,8 timeout 900 python scripts/variant_bench.py --config C2 --spp 500 --reps 5 c1:c1 default u2:u2 2>&1 | tee -a $O | cut -c1-100
timeout 600 python scripts/variant_bench.py --config C5 --spp 256 --reps 3 c1:c1 default po:po 2>&1 | tee -a $O | cut -c1-100
timeout 600 python scripts/variant_bench.py --config C1 --spp 100 --reps 5 c1:c1 default u2:u2 2>&1 | tee -a $O | cut -c1-100
timeout 600 python scripts/variant_bench.py --config C4 --spp 64 --reps 2 c1:c1 default 2>&1 | tee -a $O | cut -c1-100
ncu --set full --clock-control none --import-source on -k regex:render_wavefront -s 1 -c 1 -o gpurun_out/r2c_prof_wavefront_c2 -f python scripts/variant_bench.py --child --config C2 --spp 100 --reps 1 --mode wavefront > gpurun_out/r2c_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"
