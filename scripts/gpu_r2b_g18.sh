#!/bin/bash
# round 2, session 2, call 18: background kernel at 64 registers: tests + the bench line again + launch list at 62 spp
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu.log
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2_bench_n1.json
O=gpurun_out/r2b_g18.jsonl; : > $O
timeout 400 python scripts/variant_bench.py --config C2 --spp 62 --reps 5 default 2>&1 | tee -a $O
timeout 400 python scripts/variant_bench.py --config C2 --spp 8 --reps 5 default 2>&1 | tee -a $O
timeout 400 python scripts/variant_bench.py --config C1 --spp 100 --reps 6 default 2>&1 | tee -a $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_62spp_final.csv python scripts/variant_bench.py --child --config C2 --spp 62 --reps 2 --mode wavefront > gpurun_out/r2_launches_62_final.log 2>&1
