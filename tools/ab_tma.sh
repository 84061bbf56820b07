#!/bin/bash
# A/B timing of the column kernel variants (persistent TMA kernel on / off) on the dense and zero-padded 4096^2 loops
mkdir -p gpurun_out
for tma in 1 0; do
  echo "=== SLMGS_TMA=$tma" 
  SLMGS_TMA=$tma python tools/time_configs.py 2d 2 --dense
done
