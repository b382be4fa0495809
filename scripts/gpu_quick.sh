#!/bin/bash
# quick GPU visit: parity tests + A/B timing at reduced spp (profiling workload) + optional ncu on the pooled kernel
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|furnace|^E" gpurun_out/pytest_gpu.log | head -20
python bench.py --spp 100 --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/q_pool.json 2>gpurun_out/q_pool.err; cat gpurun_out/q_pool.json | python -c "import sys,json; d=json.load(sys.stdin); print('POOL', d['value'], d['mpaths_per_s'], d['kernel_ms_per_step'], d['roofline']['frac'])"; tail -2 gpurun_out/q_pool.err
python bench.py --spp 100 --no-cpu-baseline --steps 3 --warmup 3 --lane-per-pixel > gpurun_out/q_lane.json 2>gpurun_out/q_lane.err; cat gpurun_out/q_lane.json | python -c "import sys,json; d=json.load(sys.stdin); print('LANE', d['value'], d['mpaths_per_s'], d['kernel_ms_per_step'], d['roofline']['frac'])"
if [ "$1" = "ncu" ]; then
python bench.py --spp 50 --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:render_pool -s 1 -c 1 -o gpurun_out/prof_pool -f python bench.py --spp 50 --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log | cut -c1-300
fi
