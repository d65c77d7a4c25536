#!/usr/bin/env python
"""Stand-alone timing (CUDA events, L2 flushed) of the attention core at the preset shape, tensor-core kernels vs the
exact-fp32 bgemm + softmax kernels; also the target of the `ncu --set full -k regex:attn_` capture."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepvoice3_pytorch_b200 import ops  # noqa: E402

torch.manual_seed(0)
B, E, Td, Ts = 16, 256, 200, 128
q = torch.randn(B, E, Td, device="cuda", requires_grad=True)
k = (0.3 * torch.randn(B, E, Ts, device="cuda")).requires_grad_(True)
v = torch.randn(B, E, Ts, device="cuda", requires_grad=True)
mask = torch.zeros(B, Ts, dtype=torch.uint8, device="cuda")
mask[1, 100:] = 1
dout = torch.randn(B, E, Td, device="cuda")
dpr = 1e-3 * torch.randn(B, Td, Ts, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
reps = int(os.environ.get("ATTN_REPS", "10"))
for mode in ("tc", "fp32"):
    ops.conv_math = mode
    tf, tb = [], []
    for it in range(reps + 3):
        flush.zero_()
        s, m, e = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        s.record()
        out, probs = ops.attention_core(q, k, v, mask, 0.05, True)
        m.record()
        torch.autograd.backward([out, probs], [dout, dpr])
        e.record()
        torch.cuda.synchronize()
        if it >= 3:
            tf.append(s.elapsed_time(m) * 1e3)
            tb.append(m.elapsed_time(e) * 1e3)
    print("%-4s attention core (B=16,E=256,Td=200,Ts=128): forward %.1f us, backward %.1f us" % (
        mode, float(np.median(tf)), float(np.median(tb))), flush=True)
