"""GPU: the epilogue fusions between neighbouring convolutions (ops.Chain / Dv3TcFuse) change no arithmetic.

Forward fusion writes, from the producer's epilogue, the very bf16 planes the consumer's own split pass would have
written (same fp32 values, same dropout mask), so outputs are BIT-identical with it on or off; backward fusion runs the
producer's gate / ReLU backward in the consumer's data-gradient epilogue on the same fp32 values, so gradient planes
are bit-identical too and only the bias gradients (atomic sums in a different order) differ by round-off."""
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


def _stack(kind):
    from deepvoice3_pytorch_b200.modules import Conv1d, ConvTranspose1d, Conv1dGLU, HighwayConv1d
    torch.manual_seed(7)
    if kind == "encoder":          # 1x1+ReLU -> GLU x3 -> 1x1, like deepvoice3.Encoder.convolutions
        layers = [Conv1d(128, 256, 1, padding=0, dilation=1, std_mul=1.0), nn.ReLU(inplace=True),
                  Conv1dGLU(1, None, 256, 256, 3, dropout=0.1, dilation=1, causal=False, residual=True),
                  Conv1dGLU(1, None, 256, 256, 3, dropout=0.1, dilation=3, causal=True, residual=True),
                  Conv1dGLU(1, None, 256, 256, 3, dropout=0.1, dilation=9, causal=False, residual=False),
                  Conv1d(256, 128, 1, padding=0, dilation=1, std_mul=4.0, dropout=0.1)]
    elif kind == "converter":      # 1x1 -> up -> GLU -> GLU -> up -> GLU -> 1x1(odd width)
        layers = [Conv1d(128, 128, 1, padding=0, dilation=1, std_mul=1.0),
                  ConvTranspose1d(128, 128, 2, padding=0, stride=2, std_mul=1.0),
                  Conv1dGLU(1, None, 128, 128, 3, dropout=0.1, dilation=1, causal=False, residual=True),
                  Conv1dGLU(1, None, 128, 128, 3, dropout=0.1, dilation=3, causal=False, residual=True),
                  ConvTranspose1d(128, 128, 2, padding=0, stride=2, std_mul=4.0),
                  Conv1dGLU(1, None, 128, 128, 3, dropout=0.1, dilation=1, causal=False, residual=True),
                  Conv1d(128, 513, 1, padding=0, dilation=1, std_mul=4.0, dropout=0.1)]
    else:                          # nyanko-style highway stack with a 1x1 + ReLU head
        layers = [Conv1d(80, 256, 1, padding=0, dilation=1, std_mul=1.0), nn.ReLU(inplace=True),
                  Conv1d(256, 256, 1, padding=0, dilation=1, std_mul=2.0), nn.ReLU(inplace=True),
                  HighwayConv1d(256, 256, kernel_size=3, dilation=1, causal=True, dropout=0.1),
                  HighwayConv1d(256, 256, kernel_size=3, dilation=3, causal=True, dropout=0.1),
                  HighwayConv1d(256, 256, kernel_size=1, dilation=1, causal=False, dropout=0.1)]
    return nn.ModuleList(layers).cuda().train()


@pytest.mark.parametrize("kind,cin,T", [("encoder", 128, 200), ("converter", 128, 75), ("highway", 80, 131)])
def test_fused_epilogues_equal_separate_passes(kind, cin, T, monkeypatch):
    from deepvoice3_pytorch_b200 import ops
    from deepvoice3_pytorch_b200.modules import run_conv_stack
    monkeypatch.setattr(ops, "conv_math", "tc")
    layers = _stack(kind)
    B = 3
    gen = torch.Generator().manual_seed(1)
    x0 = torch.randn(B, cin, T, generator=gen).cuda()

    def run(fwd, bwd):
        monkeypatch.setattr(ops, "fuse_fwd", fwd)
        monkeypatch.setattr(ops, "fuse_bwd", bwd)
        ops.rng.manual_seed(4321, x0.device)
        ops.rng.start_forward()
        for p in layers.parameters():
            p.grad = None
        x = x0.clone().requires_grad_(True)
        y = run_conv_stack(layers, x)
        w = torch.cos(torch.arange(y.numel(), device="cuda", dtype=torch.float32) * 0.37).view_as(y)
        (y * w).sum().backward()
        torch.cuda.synchronize()
        return y.detach(), x.grad.clone(), {n: p.grad.clone() for n, p in layers.named_parameters()}

    y0, dx0, g0 = run(False, False)
    for fwd, bwd in ((True, False), (False, True), (True, True)):
        y1, dx1, g1 = run(fwd, bwd)
        assert torch.equal(y1, y0), "forward output changed with fuse_fwd=%s fuse_bwd=%s" % (fwd, bwd)
        assert torch.equal(dx1, dx0), "input gradient changed with fuse_fwd=%s fuse_bwd=%s" % (fwd, bwd)
        for n in g0:
            if n.endswith("bias"):                 # atomic sums in a different order
                torch.testing.assert_close(g1[n], g0[n], rtol=1e-4, atol=1e-5 * float(g0[n].abs().max() + 1), msg=n)
            else:
                assert torch.equal(g1[n], g0[n]), "%s changed with fuse_fwd=%s fuse_bwd=%s" % (n, fwd, bwd)
    assert float(dx0.abs().max()) > 0
