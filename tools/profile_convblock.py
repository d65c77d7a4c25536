#!/usr/bin/env python
"""Launch the dominant kernel a few times for `ncu --set full` (B=16, C=512, T=800, k=3 ConvBlock forward, data
gradient and weight gradient on the tensor-core path)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepvoice3_pytorch_b200 import ops  # noqa: E402

ops.conv_math = sys.argv[1] if len(sys.argv) > 1 else "tc"
B, C, T, k, d = 16, 512, 800, 3, 1
if len(sys.argv) > 6:
    B, C, T, k, d = [int(a) for a in sys.argv[2:7]]
dev = "cuda"
v = (torch.randn(2 * C, C, k, device=dev) * (4.0 / (k * C)) ** 0.5).requires_grad_(True)
g = v.detach().pow(2).sum((1, 2), keepdim=True).sqrt().requires_grad_(True)
bias = torch.zeros(2 * C, device=dev, requires_grad=True)
x = torch.randn(B, C, T, device=dev, requires_grad=True)
dy = torch.randn(B, C, T, device=dev)
for _ in range(3):
    y = ops.convblock(x, v, g, bias, None, k, d, False, ops.MODE_GLU, True)
    y.backward(dy)
torch.cuda.synchronize()
print("done")
