#!/bin/bash
# quick GPU visit: parity tests + A/B timing at reduced spp (profiling workload) + optional ncu
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|furnace|psnr|poisoned|mean radiance|^E  " gpurun_out/pytest_gpu.log | head -30
show() { python -c "import sys,json; d=json.load(open('$1')); print('$2', 'Mrays/s', round(d['value']), 'Mpaths/s', round(d['mpaths_per_s']), 'kernel_ms', round(d['kernel_ms_per_step'],2), 'frac', round(d['roofline']['frac'],4))"; }
timeout 300 python bench.py --spp 100 --no-cpu-baseline --steps 3 --warmup 3 --mode wavefront > gpurun_out/q_wf.json 2>gpurun_out/q_wf.err; show gpurun_out/q_wf.json WAVEFRONT; tail -2 gpurun_out/q_wf.err
timeout 300 python bench.py --spp 100 --no-cpu-baseline --steps 3 --warmup 3 --mode megakernel > gpurun_out/q_pool.json 2>gpurun_out/q_pool.err; show gpurun_out/q_pool.json POOL; tail -2 gpurun_out/q_pool.err
if [ "$1" = "ncu" ]; then
timeout 300 python bench.py --spp 50 --no-cpu-baseline --steps 1 --warmup 3 --mode wavefront > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wavefront -s 1 -c 1 -o gpurun_out/prof_wf -f python bench.py --spp 50 --no-cpu-baseline --steps 1 --warmup 3 --mode wavefront > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log | cut -c1-300
fi
