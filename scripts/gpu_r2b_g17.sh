#!/bin/bash
# experiments of call 16, then the round-2 evidence script
bash scripts/gpu_r2b_g16.sh
bash scripts/gpu_r2_final.sh
