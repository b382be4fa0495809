#!/bin/bash
# wavefront launch-shape sweep with the -DRTW_WF_SWEEP build (BLOCK x slots per warp)
export RTW_LIBRARY=$PWD/ray_tracing_weekend_b200/lib/librtw_cuda_sweep.so
for S in 0 1 2 3 4 5; do
RTW_WF_SHAPE=$S python bench.py --spp 100 --no-cpu-baseline --steps 3 --warmup 2 --mode wavefront 2>/dev/null | python -c "import sys,json; d=json.load(sys.stdin); print('shape $S', 'Mrays/s', round(d['value']), 'kernel_ms', round(d['kernel_ms_per_step'],2))"
done
