#!/usr/bin/env python
"""Debug aid: planes written by a producer epilogue (staging + TMA store) vs dv3_tc_split_input of the same tensor."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepvoice3_pytorch_b200 import ops  # noqa: E402
from deepvoice3_pytorch_b200.ops import lib, _p, _stream  # noqa: E402

ops.conv_math = "tc"
torch.manual_seed(0)
B, Cin, Cout, T = 4, 128, 256, 200
x = torch.randn(B, Cin, T, device="cuda")
v = torch.randn(Cout, Cin, 1, device="cuda") * (1.0 / Cin) ** 0.5
g = v.pow(2).sum((1, 2), keepdim=True).sqrt()
bias = torch.zeros(Cout, device="cuda")
with torch.enable_grad():
    xr = x.clone().requires_grad_(True)
    y = ops.conv1d(xr, v, g, bias, k=1, chain=ops.Chain(False, 0.0, False))
torch.cuda.synchronize()
pl = y._dv3_planes
ref = torch.zeros_like(pl.t)
assert pl.wg is not None, "producer did not emit the bf16 pair"
refw = torch.zeros_like(pl.wg)
lib.call("dv3_tc_split_input", _p(y.detach().contiguous()), _p(ref), 2, _p(refw), B, Cout, T, 1, 1, 0, 0.0, None, 0, _stream())
torch.cuda.synchronize()
for name, a, b in (("fp16 pair", pl.t, ref), ("bf16 pair", pl.wg, refw)):
    a16, b16 = a.view(torch.int16), b.view(torch.int16)
    bad = (a16 != b16)
    print(name, "mismatch fraction per plane:", [float(bad[i].float().mean()) for i in range(2)])
    if bad.any():
        idx = bad[0].nonzero()[:8]
        print(" first mismatches (b,t,c):", idx.tolist())
        bt = bad[0, 0]                      # (T, C) of batch 0
        print(" rows with any mismatch (first 16):", bt.any(1).nonzero().flatten()[:16].tolist())
        print(" cols with any mismatch (first 40):", bt.any(0).nonzero().flatten()[:40].tolist())
        # where did the value of (t=1, c=8..15) land?
        want = b[0, 0, 1, 8:16]
        hits = (a[0, 0].unsqueeze(-1) == want[0]).any(-1).nonzero()[:6]
        print(" value ref[0,0,1,8] found in fused planes at (t,c):", hits.tolist())
