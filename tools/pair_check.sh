#!/bin/bash
# CTA-pair kernels: block-level parity on a few shapes (each in its own process, short timeout), then timing
export DV3_TC_PAIR=2
for c in 2 5 7 8; do
  timeout 60 python tools/tc_debug.py $c 2>&1 | tail -1 | cut -c1-220
  echo "rc=$?"
done
DV3_OVERLAP_WGRAD=0 timeout 120 python tools/tc_time.py 2>&1 | tail -5 | cut -c1-300
