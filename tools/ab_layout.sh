#!/bin/bash
# A/B timing: field layout (row-major / row-pair interleaved) x column tile width on the dense and zero-padded 4096^2 loops
for cfg in "0 0" "1 0" "1 512" "0 512"; do
  set -- $cfg
  echo "=== SLMGS_PAIRS=$1 SLMGS_COL_THREADS=$2"
  SLMGS_PAIRS=$1 SLMGS_COL_THREADS=$2 python tools/time_configs.py 2d 2 --dense
done
