#!/bin/bash
export DV3_OVERLAP_WGRAD=0 TC_TIME_FIRST=1
for dbg in 5 4 6; do
  echo "== DV3_TC_DEBUG=$dbg (bit0 no epilogue, bit1 no MMA, bit2 no TMA)"
  DV3_TC_DEBUG=$dbg timeout 100 python tools/tc_time.py 2>&1 | tail -1 | cut -c1-200
done
