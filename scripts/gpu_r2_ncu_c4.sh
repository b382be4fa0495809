#!/bin/bash
# ncu --set full of the wavefront kernel on C4 (1 M spheres, 49 988 lights) at 8 spp: CONNECT stage
mkdir -p gpurun_out
python scripts/variant_bench.py --config C4 --spp 8 --reps 1 connect > gpurun_out/plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:render_wavefront -s 1 -c 1 -o gpurun_out/r2_prof_wavefront_c4_connect -f python scripts/variant_bench.py --child --config C4 --spp 8 --reps 1 --mode wavefront > gpurun_out/ncu_c4_connect.log 2>&1
echo "ncu connect rc=$?"
