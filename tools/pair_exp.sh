#!/bin/bash
# where does the time of the persistent tensor-core kernels go?  (timing experiments; results with DV3_TC_DEBUG != 0 are wrong on purpose)
export DV3_OVERLAP_WGRAD=0 TC_TIME_FIRST=1
for pair in 0 2; do
  for dbg in 0 1 2 3; do
    echo "== DV3_TC_PAIR=$pair DV3_TC_DEBUG=$dbg"
    DV3_TC_PAIR=$pair DV3_TC_DEBUG=$dbg timeout 100 python tools/tc_time.py 2>&1 | tail -1 | cut -c1-200
  done
done
echo "== taps, 2 stages"; DV3_TC_MAXSTAGES=2 timeout 100 python tools/tc_time.py 2>&1 | tail -1 | cut -c1-200
echo "== pair, 2 stages"; DV3_TC_PAIR=2 DV3_TC_MAXSTAGES=2 timeout 100 python tools/tc_time.py 2>&1 | tail -1 | cut -c1-200
