#!/bin/bash
# Round-end evidence run: everything the profiles/ directory cites, from the current HEAD.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_n1.json
python bench.py --mode megakernel --no-cpu-baseline > gpurun_out/bench_n1_megakernel.json 2>/dev/null
python bench.py --mode megakernel --lane-per-pixel --no-cpu-baseline > gpurun_out/bench_n1_lane.json 2>/dev/null
nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/l2_bandwidth.cu -o /tmp/l2bw && /tmp/l2bw > gpurun_out/l2_bandwidth.json; cat gpurun_out/l2_bandwidth.json
timeout 600 python scripts/stress_configs.py C1 C5 C4 C3 > gpurun_out/stress.jsonl 2>&1; tail -2 gpurun_out/stress.jsonl | cut -c1-200
python bench.py > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python bench.py > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:render_wavefront -s 4 -c 1 -o gpurun_out/prof_wavefront_r1 -f python bench.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
python scripts/stress_configs.py C4 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:render_wavefront -s 1 -c 1 -o gpurun_out/prof_wavefront_c4 -f python scripts/stress_configs.py C4 > gpurun_out/ncu_c4.log 2>&1
echo "ncu c4 rc=$?"
