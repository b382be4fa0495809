#!/bin/bash
# strong scaling on one multi-GPU box, launched exactly as the driver does; N = number of GPUs of the box (8 or 4), pixel partition (default) then sample partition
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "N=$N pixels rc=$? lines=$(wc -l < gpurun_out/r2_bench_n$N.json)"
RTW_MULTI_PARTITION=samples python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+N)) bench.py --gpus $N --steps 4 --warmup 3 --no-c3 > gpurun_out/r2_bench_n${N}_samples.json 2> gpurun_out/r2_bench_n${N}_samples.err; echo "N=$N samples rc=$?"
for f in gpurun_out/r2_bench_n$N.json gpurun_out/r2_bench_n${N}_samples.json; do python -c "
import json
d=json.loads(open('$f').read())
print('$f', round(d['value']), 'Mrays/s', round(d['ms_per_step'],2), 'ms/step kernel', round(d['kernel_ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'c3', (d.get('c3') or {}).get('ms_per_step'))
"; done
