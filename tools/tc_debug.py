#!/usr/bin/env python
"""Stage-by-stage check of the tensor-core ConvBlock path against the exact-fp32 kernels (GPU only)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepvoice3_pytorch_b200 import ops  # noqa: E402


TC_MATH = os.environ.get("TC_MATH", "tc")


def run_plain(B, Cin, Cout, T, relu, convt=False):
    dev = "cuda"
    torch.manual_seed(B + Cin + Cout + T)
    if convt:
        v = torch.randn(Cin, Cout, 2, device=dev) * (1.0 / (2 * Cin)) ** 0.5
        g = v.pow(2).sum((1, 2), keepdim=True).sqrt() * (1 + 0.2 * torch.randn(Cin, 1, 1, device=dev))
        Co = Cout
    else:
        v = torch.randn(Cout, Cin, 1, device=dev) * (1.0 / Cin) ** 0.5
        g = v.pow(2).sum((1, 2), keepdim=True).sqrt() * (1 + 0.2 * torch.randn(Cout, 1, 1, device=dev))
        Co = Cout
    bias = 0.1 * torch.randn(Co, device=dev)
    x = torch.randn(B, Cin, T, device=dev)
    dy = torch.randn(B, Co, 2 * T if convt else T, device=dev)
    res = {}
    for math in ("fp32", TC_MATH):
        ops.conv_math = math
        leaves = [t.clone().requires_grad_(True) for t in (x, v, g, bias)]
        if convt:
            y = ops.conv_transpose1d_k2s2(*leaves)
        else:
            y = ops.conv1d(leaves[0], leaves[1], leaves[2], leaves[3], relu=relu)
        y.backward(dy)
        torch.cuda.synchronize()
        res[math] = [y.detach()] + [t.grad for t in leaves]
    names = ["y", "dx", "dv", "dg", "dbias"]
    out = " ".join("%s=%.2e" % (n, rel(a, b)) for n, a, b in zip(names, res[TC_MATH], res["fp32"]))
    print("%s B=%d Cin=%d Cout=%d T=%d relu=%d : %s" % ("convT" if convt else "conv1x1", B, Cin, Cout, T, relu, out),
          flush=True)


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def run(B, C, T, k, d, causal, mode, residual, p_drop=0.0):
    dev = "cuda"
    torch.manual_seed(B + C + T + d)
    v = torch.randn(2 * C, C, k, device=dev) * (4.0 / (k * C)) ** 0.5
    g = v.pow(2).sum((1, 2), keepdim=True).sqrt() * (1 + 0.2 * torch.randn(2 * C, 1, 1, device=dev))
    bias = 0.1 * torch.randn(2 * C, device=dev)
    x = torch.randn(B, C, T, device=dev)
    dy = torch.randn(B, C, T, device=dev)
    res = {}
    for math in ("fp32", TC_MATH):
        ops.conv_math = math
        ops.rng.manual_seed(99, x.device)
        ops.rng.start_forward()
        leaves = [t.clone().requires_grad_(True) for t in (x, v, g, bias)]
        y = ops.convblock(leaves[0], leaves[1], leaves[2], leaves[3], None, k, d, causal, mode, residual,
                          p_drop=p_drop, training=p_drop > 0)
        torch.cuda.synchronize()
        y.backward(dy)
        torch.cuda.synchronize()
        res[math] = [y.detach()] + [t.grad for t in leaves]
    names = ["y", "dx", "dv", "dg", "dbias"]
    out = " ".join("%s=%.2e" % (n, rel(a, b)) for n, a, b in zip(names, res[TC_MATH], res["fp32"]))
    print("B=%d C=%d T=%d k=%d d=%d causal=%d mode=%d res=%d p=%.2f : %s" % (B, C, T, k, d, causal, mode, residual,
                                                                           p_drop, out), flush=True)


CASES = [
    (2, 128, 64, 1, 1, False, 0, True, 0.0),     # 0: known good
    (2, 256, 128, 1, 1, False, 0, True, 0.0),    # 1: stage reuse, no negative coords
    (2, 128, 128, 3, 1, False, 0, True, 0.0),    # 2: negative coords
    (2, 128, 128, 2, 1, True, 0, True, 0.0),     # 3: causal k=2: fwd offsets -1,0 ; dgrad +1,0
    (2, 256, 200, 3, 9, True, 0, False, 0.0),
    (4, 512, 128, 3, 27, False, 0, True, 0.0),
    (2, 256, 200, 3, 3, True, 1, True, 0.0),
    (3, 256, 800, 3, 3, False, 0, True, 0.05),
    (16, 512, 800, 3, 1, False, 0, True, 0.0),
]

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "fwd":
        # forward only, for fault isolation
        B, C, T, k, d, causal, mode, residual, p = CASES[int(sys.argv[2])]
        ops.conv_math = "bf16x3"
        v = torch.randn(2 * C, C, k, device="cuda") * 0.05
        g = v.pow(2).sum((1, 2), keepdim=True).sqrt()
        with torch.no_grad():
            y = ops.convblock(torch.randn(B, C, T, device="cuda"), v, g, torch.zeros(2 * C, device="cuda"), None, k,
                              d, causal, mode, residual)
        torch.cuda.synchronize()
        print("fwd-only case", sys.argv[2], "ok", float(y.abs().mean()))
    elif sys.argv[1] == "plain":
        run_plain(16, 256, 512, 128, True)
        run_plain(16, 80, 256, 200, True)
        run_plain(16, 256, 80, 200, False)
        run_plain(4, 512, 513, 800, False)
        run_plain(2, 513, 513, 800, True)
        run_plain(16, 256, 256, 200, False, convt=True)
        run_plain(4, 512, 512, 400, False, convt=True)
    else:
        run(*CASES[int(sys.argv[1])])
