#!/bin/bash
# ncu --set full of the wavefront kernel on C2 (1080p, 100 spp profiling workload)
mkdir -p gpurun_out
python scripts/variant_bench.py --config C2 --spp 100 --reps 1 default > gpurun_out/plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:render_wavefront -s 1 -c 1 -o gpurun_out/r2_prof_wavefront_c2 -f python scripts/variant_bench.py --child --config C2 --spp 100 --reps 1 --mode wavefront > gpurun_out/ncu_c2.log 2>&1
echo "ncu c2 rc=$?"
