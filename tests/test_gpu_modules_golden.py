"""GPU parity of the package's MODULE classes (the objects a reference user touches) against the golden vectors the
live reference produced for the same classes: multi-speaker Conv1dGLU, ConvTranspose1d, Linear, AttentionLayer
(+/- projections, +/- memory mask) and SinusoidalEncoding incl. per-utterance rates.  State dicts load key-for-key;
outputs and every gradient at north_star's rtol=1e-3 / atol=1e-4, position tables / integer lookups exactly."""
import numpy as np
import pytest
import torch

import golden_util as G
from test_gpu_blocks import close, grad_close

pytestmark = pytest.mark.gpu
BLOCKS = G.load("blocks.npz")


def _run_module(case, module, call):
    dev = "cuda"
    res = module.load_state_dict(G.tensors(case["sd"]), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    module = module.to(dev).train()                       # dropout = 0 in every fixture
    ins = G.tensors(case["in"], dev)
    for v in ins.values():
        if v.is_floating_point():
            v.requires_grad_(True)
    outs = call(module, ins)
    outs = outs if isinstance(outs, tuple) else (outs,)
    for i, o in enumerate(outs):
        close(o, case["out"][str(i)], what="out%d" % i)
    sum((o * G.loss_weights(o.shape, i, dev)).sum() for i, o in enumerate(outs)).backward()
    params = dict(module.named_parameters())
    for k, ref in case.get("gsd", {}).items():
        assert params[k].grad is not None, k
        grad_close(params[k].grad, ref, what="grad " + k)
    for k, ref in case.get("gin", {}).items():
        grad_close(ins[k].grad, ref, what="grad in " + k)


@pytest.mark.parametrize("math", ["fp32", "tc"])
@pytest.mark.parametrize("name", ["glu_spk0", "glu_spk1"])
def test_multispeaker_conv1dglu_module(name, math, monkeypatch):
    """reference modules.py:112-167 with the softsign(Linear(speaker_embed)) bias on the `a` half."""
    from deepvoice3_pytorch_b200 import ops
    from deepvoice3_pytorch_b200.modules import Conv1dGLU
    monkeypatch.setattr(ops, "conv_math", math)
    case = BLOCKS[name]
    m = {k: G.meta_scalar(case, k) for k in ("k", "d", "causal", "residual")}
    mod = Conv1dGLU(4, 16, 32, 32, int(m["k"]), dropout=0.0, dilation=int(m["d"]), causal=bool(m["causal"]),
                    residual=bool(m["residual"]))
    _run_module(case, mod, lambda f, L: f(L["x"], L["spk"]))


@pytest.mark.parametrize("math", ["fp32", "tc"])
def test_conv_transpose_module(math, monkeypatch):
    from deepvoice3_pytorch_b200 import ops
    from deepvoice3_pytorch_b200.modules import ConvTranspose1d
    monkeypatch.setattr(ops, "conv_math", math)
    _run_module(BLOCKS["convT"], ConvTranspose1d(32, 48, 2, padding=0, stride=2), lambda f, L: f(L["x"]))


@pytest.mark.parametrize("math", ["fp32", "tc"])
def test_linear_module(math, monkeypatch):
    from deepvoice3_pytorch_b200 import ops
    from deepvoice3_pytorch_b200.modules import Linear
    monkeypatch.setattr(ops, "conv_math", math)
    _run_module(BLOCKS["linear"], Linear(16, 64), lambda f, L: f(L["x"]))


@pytest.mark.parametrize("math", ["fp32", "tc"])
@pytest.mark.parametrize("name,kp,vp", [("attn0", True, True), ("attn1", False, False), ("attn2", True, False)])
def test_attention_layer_module(name, kp, vp, math, monkeypatch):
    """reference deepvoice3.py:132-176 through the reference call signature (query (B,Td,C), (keys (B,E,Ts),
    values (B,Ts,E)), mask)."""
    from deepvoice3_pytorch_b200 import ops
    from deepvoice3_pytorch_b200.deepvoice3 import AttentionLayer
    monkeypatch.setattr(ops, "conv_math", math)
    case = BLOCKS[name]
    lengths = case["meta"]["lengths"]
    Ts = case["in"]["keys"].shape[-1]
    mask = None
    if lengths.size:
        mask = ~(torch.arange(Ts)[None, :] < torch.as_tensor(lengths)[:, None])
        mask = mask.cuda()
    mod = AttentionLayer(48, 32, dropout=0.0, key_projection=kp, value_projection=vp)
    _run_module(case, mod, lambda f, L: f(L["query"], (L["keys"], L["values"]), mask=mask))


@pytest.mark.parametrize("name", [n for n in BLOCKS if n.startswith("sin") and n != "sin_batch"])
def test_sinusoidal_encoding_module(name):
    """reference modules.py:34-64 / tests/test_embedding.py: scalar rate; padding position 0 -> zero row."""
    from deepvoice3_pytorch_b200.modules import SinusoidalEncoding, position_encoding_init
    case = BLOCKS[name]
    w = float(case["meta"]["w"])
    n, d = case["out"]["table"].shape
    mod = SinusoidalEncoding(n, d).cuda()
    x = torch.from_numpy(case["in"]["x"]).cuda()
    out = mod(x, w)
    close(out, case["out"]["0"], rtol=1e-5, atol=2e-6, what=name)
    assert float(out[1, 100:].abs().max()) == 0.0           # padding positions are exactly zero
    np.testing.assert_array_equal(position_encoding_init(n, d, position_rate=w).numpy(), case["out"]["table"])


def test_sinusoidal_encoding_per_utterance_rates():
    """multi-speaker path (reference modules.py:57-64): one position rate per batch row."""
    from deepvoice3_pytorch_b200.modules import SinusoidalEncoding
    case = BLOCKS["sin_batch"]
    table = torch.from_numpy(case["sd"]["weight"])
    mod = SinusoidalEncoding(*table.shape)
    mod.load_state_dict({"weight": table})
    mod = mod.cuda()
    out = mod(torch.from_numpy(case["in"]["x"]).cuda(), torch.from_numpy(case["meta"]["w"]).cuda())
    close(out, case["out"]["0"], rtol=1e-5, atol=2e-6)


def test_attention_window_last_attended():
    """reference deepvoice3.py:150-156: with last_attended = n only keys in [n - window_backward, n + window_ahead) get
    probability mass, in the batch forward too; result equals the oracle's masked softmax."""
    from deepvoice3_pytorch_b200.deepvoice3 import AttentionLayer
    case = BLOCKS["attn1"]
    mod = AttentionLayer(48, 32, dropout=0.0, key_projection=False, value_projection=False, window_ahead=3,
                         window_backward=1)
    mod.load_state_dict(G.tensors(case["sd"]))
    mod = mod.cuda().eval()
    ins = G.tensors(case["in"], "cuda")
    with torch.no_grad():
        _, probs = mod(ins["query"], (ins["keys"], ins["values"]), mask=None, last_attended=5)
        _, full = mod(ins["query"], (ins["keys"], ins["values"]), mask=None)
    assert float(probs[:, :, :4].abs().max()) == 0.0 and float(probs[:, :, 8:].abs().max()) == 0.0
    want = full[:, :, 4:8] / full[:, :, 4:8].sum(-1, keepdim=True)
    close(probs[:, :, 4:8], want, rtol=1e-4, atol=1e-6)
