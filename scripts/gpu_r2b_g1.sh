#!/bin/bash
# round 2, session 2, call 1: GPU tests at HEAD + 256-bit node loads + 128-bit general-scene record loads; Philox out of line on the sphere
# kernels (i-cache); fresh ncu capture of C2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_g1_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_g1_pytest.log
O=gpurun_out/r2b_g1.jsonl; : > $O
timeout 300 python scripts/variant_bench.py --config C2 --spp 100 --reps 5 default philox:philox noldg256:noldg256 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C4 --spp 16 --reps 3 default noldg256:noldg256 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C5 --spp 64 --reps 3 default noldg256:noldg256 philox:philox 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config C1 --spp 100 --reps 5 default philox:philox 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config cornell_box --spp 256 --reps 3 default gscalar:gscalar 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config debugging_scene --spp 64 --reps 3 default gscalar:gscalar 2>&1 | tee -a $O
timeout 300 python scripts/variant_bench.py --config checkered_spheres --spp 64 --reps 3 default gscalar:gscalar 2>&1 | tee -a $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wavefront -s 1 -c 1 -o gpurun_out/r2b_prof_wavefront_c2 -f python scripts/variant_bench.py --child --config C2 --spp 100 --reps 1 --mode wavefront > gpurun_out/r2b_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"
