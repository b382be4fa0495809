#!/bin/bash
# round 2, session 2, call 3: where the per-launch fixed cost sits now (launch lists at 8 / 62 spp), ncu capture of the headline kernel at HEAD
mkdir -p gpurun_out
for SPP in 8 62; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2b_launches_${SPP}spp.csv python scripts/variant_bench.py --child --config C2 --spp $SPP --reps 3 --mode wavefront > gpurun_out/r2b_launches_${SPP}.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wavefront -s 1 -c 1 -o gpurun_out/r2b_prof_wavefront_c2_selfhit -f python scripts/variant_bench.py --child --config C2 --spp 100 --reps 1 --mode wavefront > gpurun_out/r2b_ncu_c2_selfhit.log 2>&1
echo "ncu c2 rc=$?"
