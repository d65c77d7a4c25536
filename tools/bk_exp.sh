#!/bin/bash
# SWIZZLE_64B (BK=32) vs SWIZZLE_128B (BK=64) operands in the persistent gated kernel: parity, full time, MMA-only time
export DV3_OVERLAP_WGRAD=0 TC_TIME_FIRST=1 DV3_TC_TAPS=0
DV3_TC_PERSIST_BK=64 timeout 60 python tools/tc_debug.py 8 2>&1 | tail -1 | cut -c1-200
for bk in 32 64; do
  for dbg in 0 5; do
    echo "== BK=$bk DV3_TC_DEBUG=$dbg"
    DV3_TC_PERSIST_BK=$bk DV3_TC_DEBUG=$dbg timeout 100 python tools/tc_time.py 2>&1 | tail -1 | cut -c1-200
  done
done
