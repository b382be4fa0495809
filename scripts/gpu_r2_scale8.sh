#!/bin/bash
# strong-scaling on one 8-GPU box, launched exactly as the driver does: N = 8 then 4 (N = 1, 2 are measured on their own boxes)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "N=$N rc=$? lines=$(wc -l < gpurun_out/r2_bench_n$N.json)"
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read())
print('N=$N', round(d['value']), 'Mrays/s', round(d['ms_per_step'],2), 'ms/step kernel', round(d['kernel_ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['config']['parallelism'], 'c3', d.get('c3',{}).get('ms_per_step'))
"
done
