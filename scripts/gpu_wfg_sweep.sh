#!/bin/bash
# launch-shape sweep of the general-scene wavefront (cornell_box) with the -DRTW_WF_SWEEP build
export RTW_LIBRARY=$PWD/ray_tracing_weekend_b200/lib/librtw_cuda_sweep.so
for S in 0 1 2 3 4 5; do echo "shape $S: $(RTW_WFG_SHAPE=$S python scripts/general_renderers.py 2>&1 | head -1)"; done
