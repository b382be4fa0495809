#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (own + reference arm), ncu launch list + full capture.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json | cut -c1-400
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
python bench.py --mode megakernel --no-cpu-baseline > gpurun_out/bench_n1_megakernel.json 2> gpurun_out/bench_n1_mk.err; cat gpurun_out/bench_n1_megakernel.json | cut -c1-300
python bench.py --mode megakernel --lane-per-pixel --no-cpu-baseline > gpurun_out/bench_n1_lane.json 2> gpurun_out/bench_n1_lane.err; cat gpurun_out/bench_n1_lane.json | cut -c1-300
if [ "$1" = "ncu" ]; then
python bench.py > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python bench.py > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:render_wavefront -s 4 -c 1 -o gpurun_out/prof_wavefront_r1 -f python bench.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log | cut -c1-200
fi
