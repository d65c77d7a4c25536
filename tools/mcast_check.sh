#!/bin/bash
# weight-multicast cluster mode of the tap-reuse kernel: parity, then timing / feed-only decomposition
for c in 2 5 8; do
  DV3_TC_MCAST=1 DV3_TC_TAPS=2 timeout 60 python tools/tc_debug.py $c 2>&1 | tail -1 | cut -c1-220
  echo "rc=$?"
done
export DV3_OVERLAP_WGRAD=0
for dbg in 0 3; do
  echo "== DV3_TC_MCAST=1 DV3_TC_DEBUG=$dbg"
  DV3_TC_MCAST=1 DV3_TC_DEBUG=$dbg TC_TIME_FIRST=3 timeout 100 python tools/tc_time.py 2>&1 | tail -3 | cut -c1-200
done
echo "== bench MCAST=1"; DV3_TC_MCAST=1 timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
echo "== bench MCAST=0"; timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
