#!/bin/bash
# Round-2 evidence at the round's last commit (1 GPU): launch-shape sweep first (tuning build), then tests, smoke, both bench arms, stress configs,
# general scenes, launch list of the bench command, ncu --set full of the wavefront kernel inside bench.py
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2>/dev/null; wc -l gpurun_out/r2_bench_ref.json
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; wc -l gpurun_out/r2_bench_n1.json; cut -c1-200 gpurun_out/r2_bench_n1.json
python bench.py --mode megakernel --no-cpu-baseline --no-f64 --no-c3 > gpurun_out/r2_bench_n1_mega.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_n1_mega.json
timeout 900 python scripts/stress_configs.py > gpurun_out/r2_stress_configs.jsonl 2> gpurun_out/r2_stress_configs.err; echo "stress rc=$?"
timeout 600 python scripts/general_configs.py > gpurun_out/r2_general_scenes.jsonl 2> gpurun_out/r2_general_scenes.err; echo "general rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches_bench_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-f64 --no-c3 > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:render_wavefront -s 2 -c 1 -o gpurun_out/r2_prof_wavefront_bench -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-f64 --no-c3 > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"
