#!/usr/bin/env python
"""Time the three tensor-core kernels of the big ConvBlock shapes in isolation (CUDA events, no L2 flush)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepvoice3_pytorch_b200 import ops  # noqa: E402

ops.conv_math = "tc"
dev = "cuda"
SHAPES = [(16, 512, 800, 3, 1), (16, 512, 800, 3, 27), (16, 256, 800, 3, 3), (16, 512, 128, 3, 9), (16, 256, 200, 3, 9)]
if os.environ.get("TC_TIME_FIRST"):
    SHAPES = SHAPES[:int(os.environ["TC_TIME_FIRST"])]
for (B, C, T, k, d) in SHAPES:
    v = (torch.randn(2 * C, C, k, device=dev) * (4.0 / (k * C)) ** 0.5).requires_grad_(True)
    g = v.detach().pow(2).sum((1, 2), keepdim=True).sqrt().requires_grad_(True)
    bias = torch.zeros(2 * C, device=dev, requires_grad=True)
    x = torch.randn(B, C, T, device=dev, requires_grad=True)
    dy = torch.randn(B, C, T, device=dev)
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            y = ops.convblock(x, v, g, bias, None, k, d, False, ops.MODE_GLU, True)
            y.backward(dy)
        torch.cuda.synchronize()
    rows = {}
    for e in prof.key_averages():
        if "tc_conv_kernel" in e.key:
            rows[e.key.split("tc_conv_kernel")[1][:18]] = e.device_time_total / e.count
        elif "tc_conv_pair64" in e.key:
            rows["pair64" + e.key.split("tc_conv_pair64_kernel")[1][:4]] = e.device_time_total / e.count
        elif "tc_conv_pair" in e.key:
            rows["pair" + e.key.split("tc_conv_pair_kernel")[1][:4]] = e.device_time_total / e.count
        elif "tc_conv_taps" in e.key:
            rows["taps" + e.key.split("tc_conv_taps_kernel")[1][:10]] = e.device_time_total / e.count
        elif "tc_conv_persist" in e.key:
            rows["persist" + e.key.split("tc_conv_persist_kernel")[1][:10]] = e.device_time_total / e.count
        elif "tc_wgrad_mn" in e.key or "wn_bwd" in e.key:
            rows[e.key.split("dv3::")[-1][:18]] = e.device_time_total / e.count
    flops = 2.0 * B * T * 2 * C * C * k
    print("B=%d C=%d T=%d: " % (B, C, T) + "  ".join("%s %.1f us (%.0f TF)" % (kk, vv, flops / vv / 1e6)
                                                      for kk, vv in sorted(rows.items())), flush=True)
