#!/bin/bash
for L in 1 2 3 4 6 8; do
RTW_BVH_MAX_LEAF=$L python bench.py --spp 100 --no-cpu-baseline --steps 3 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.load(sys.stdin); e=d['events_per_step']; print('leaf $L', 'Mrays/s', round(d['value']), 'kernel_ms', round(d['kernel_ms_per_step'],2), 'nodes/ray', round(e['node_visits']/e['rays'],2), 'sph/ray', round(e['sphere_tests']/e['rays'],2))"
done
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
