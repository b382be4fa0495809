#!/bin/bash
# round 2, session 3, call 6: new default (Philox out of line, merged Lambertian halves in the flat-list kernel): tests, all configs, further code-size tickets, ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_g6_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c_g6_pytest.log
O=gpurun_out/r2c_g6.jsonl; : > $O
timeout 900 python scripts/variant_bench.py --config C2 --spp 500 --reps 3 default cr:cr u1:u1 u2:u2 cr_u2:cr_u2 2>&1 | tee -a $O | cut -c1-100
RTW_DEBUG_OWN=0,8 timeout 900 python scripts/variant_bench.py --config C2 --spp 500 --reps 5 default t4k,RTW_CHEAP_TAIL_PATHS=4096 2>&1 | tee -a $O | cut -c1-100
timeout 600 python scripts/variant_bench.py --config C5 --spp 256 --reps 3 default cr:cr 2>&1 | tee -a $O | cut -c1-100
timeout 600 python scripts/variant_bench.py --config C1 --spp 100 --reps 5 default cr:cr 2>&1 | tee -a $O | cut -c1-100
timeout 600 python scripts/variant_bench.py --config C4 --spp 64 --reps 2 default 2>&1 | tee -a $O | cut -c1-100
timeout 600 python scripts/variant_bench.py --config C2 --spp 100 --reps 3 --mode megakernel default 2>&1 | tee -a $O | cut -c1-100
ncu --set full --clock-control none --import-source on -k regex:render_wavefront -s 1 -c 1 -o gpurun_out/r2c_prof_wavefront_c2_b -f python scripts/variant_bench.py --child --config C2 --spp 100 --reps 1 --mode wavefront > gpurun_out/r2c_ncu_c2_b.log 2>&1
echo "ncu c2 rc=$?"
