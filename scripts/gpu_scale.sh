#!/bin/bash
# strong-scaling sweep on one 8-GPU box, launched exactly as the driver does
mkdir -p gpurun_out
for N in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "N=$N rc=$?"
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1]
print('N=$N', round(d['value']), 'Mrays/s', round(d['mpaths_per_s']), 'Mpaths/s', round(d['ms_per_step'],2), 'ms/step kernel', round(d['kernel_ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'other', d['other_renderer'])
"
done
CUDA_VISIBLE_DEVICES=0 python bench.py --no-cpu-baseline > gpurun_out/bench_n1_same_box.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/bench_n1_same_box.json')); print('N=1', round(d['value']), 'Mrays/s', round(d['ms_per_step'],2), 'ms/step')"
