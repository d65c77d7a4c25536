#!/bin/bash
# one-shot check of the tap-reuse kernels: block-level parity (both descriptor variants), then tests + timing
for bo in 0 1; do
  echo "== DV3_TC_TAPS_BASEOFF=$bo"
  for c in 2 3 4 5 6 7 8; do
    DV3_TC_TAPS_BASEOFF=$bo timeout 120 python tools/tc_debug.py $c 2>&1 | tail -1 | cut -c1-200
  done
done
