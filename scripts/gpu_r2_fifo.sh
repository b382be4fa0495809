#!/bin/bash
# round 2: FIFO (ring) work lists vs LIFO (stack) lists; two-level candidate pre-pass
mkdir -p gpurun_out
O=gpurun_out/r2_fifo.jsonl; : > $O
for SPP in 8 62; do
python scripts/variant_bench.py --config C2 --spp $SPP --reps 6 fifo lifo:lifo 2>&1 | tee -a $O
done
python scripts/variant_bench.py --config C2 --spp 250 --reps 3 fifo lifo:lifo 2>&1 | tee -a $O
python scripts/variant_bench.py --config C1 --spp 100 --reps 6 fifo lifo:lifo 2>&1 | tee -a $O
python scripts/variant_bench.py --config C5 --spp 64 --reps 3 fifo lifo:lifo 2>&1 | tee -a $O
python scripts/variant_bench.py --config C4 --spp 16 --reps 2 fifo lifo:lifo 2>&1 | tee -a $O
python scripts/fixed_cost_probe.py 2>&1 | tee gpurun_out/r2_fixed_cost_probe_fifo.jsonl
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_fifo_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_fifo_pytest.log
