#!/bin/bash
# Round-2 first item: validate + time the BK=64 CTA-pair kernel (tc_conv_pair64_kernel, opt-in DV3_TC_PAIR64).
# Each parity case runs in its own process under a short timeout (a protocol bug would hang the kernel).
for c in 2 5 8; do
  DV3_TC_PAIR64=2 timeout 60 python tools/tc_debug.py $c 2>&1 | tail -1 | cut -c1-220
  echo "rc=$?"
done
export DV3_OVERLAP_WGRAD=0
for dbg in 0 5; do
  echo "== DV3_TC_PAIR64=1 DV3_TC_DEBUG=$dbg (5 = MMAs only)"
  DV3_TC_PAIR64=1 DV3_TC_DEBUG=$dbg TC_TIME_FIRST=3 timeout 100 python tools/tc_time.py 2>&1 | tail -3 | cut -c1-220
done
echo "== bench PAIR64=1"; DV3_TC_PAIR64=1 timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
echo "== bench default"; timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
