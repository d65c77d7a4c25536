#!/bin/bash
export DV3_TC_TAPS=0 DV3_TC_PERSIST_BK=64 DV3_TC_SMALL_BK=64
for c in 2 4 5 6 7 8; do timeout 60 python tools/tc_debug.py $c 2>&1 | tail -1 | cut -c1-200; done
timeout 100 python tools/tc_debug.py plain 2>&1 | tail -7 | cut -c1-160
DV3_OVERLAP_WGRAD=0 timeout 100 python tools/tc_time.py 2>&1 | tail -5 | cut -c1-250
echo "== bench bk64"; timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
unset DV3_TC_TAPS DV3_TC_PERSIST_BK DV3_TC_SMALL_BK
echo "== bench default"; timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
echo "== bench persist64 only (taps off)"; DV3_TC_TAPS=0 DV3_TC_PERSIST_BK=64 timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
