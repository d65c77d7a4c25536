#!/usr/bin/env python
"""Micro-benchmark of the ConvBlock kernels at the BASELINE.md canonical shapes (CUDA events, L2 flushed
between iterations).  Prints one JSON line per shape; not the headline bench (see bench.py)."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepvoice3_pytorch_b200 import ops  # noqa: E402


def timeit(fn, iters=20, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def main():
    dev = "cuda"
    if len(sys.argv) > 1:
        ops.conv_math = sys.argv[1]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    shapes = [(2, 64, 128, 3, 1, False), (16, 256, 200, 3, 9, True), (16, 512, 128, 3, 9, False),
              (16, 256, 800, 3, 3, False), (16, 512, 800, 3, 1, False)]
    for (B, C, T, k, d, causal) in shapes:
        v = (torch.randn(2 * C, C, k, device=dev) * (4.0 / (k * C)) ** 0.5).requires_grad_(True)
        g = v.detach().pow(2).sum((1, 2), keepdim=True).sqrt().requires_grad_(True)
        bias = torch.zeros(2 * C, device=dev, requires_grad=True)
        x = torch.randn(B, C, T, device=dev, requires_grad=True)
        dy = torch.randn(B, C, T, device=dev)

        def fwd():
            with torch.no_grad():
                return ops.convblock(x, v, g, bias, None, k, d, causal, ops.MODE_GLU, True)

        def fwdbwd():
            y = ops.convblock(x, v, g, bias, None, k, d, causal, ops.MODE_GLU, True)
            y.backward(dy)

        tf = timeit(fwd, flush=flush)
        tfb = timeit(fwdbwd, flush=flush)
        flops = 2.0 * B * T * 2 * C * C * k
        bytes_fwd = 4.0 * (2 * B * C * T + 2 * C * C * k + 4 * C)
        print(json.dumps(dict(math=ops.conv_math, shape=[B, C, T, k, d], fwd_us=tf * 1e6, fwdbwd_us=tfb * 1e6,
                              fwd_tflops=flops / tf / 1e12, fwdbwd_tflops=3 * flops / tfb / 1e12,
                              fwd_alg_GBps=bytes_fwd / tf / 1e9)))


if __name__ == "__main__":
    main()
