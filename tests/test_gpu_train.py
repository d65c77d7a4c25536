"""GPU: the training-step machinery (flat optimizer kernels, device-side losses, graph replay)."""
import numpy as np
import pytest
import torch


def adam_clip_reference(p, g, m, v, t, lr, betas, eps, max_norm, grad_scale=1.0):
    """clip_grad_norm_ (torch/nn/utils/clip_grad.py) followed by torch.optim.Adam's update, in torch ops."""
    g = g * grad_scale
    total = g.norm()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0) if max_norm > 0 else 1.0
    g = g * coef
    m = betas[0] * m + (1 - betas[0]) * g
    v = betas[1] * v + (1 - betas[1]) * g * g
    bc1, bc2 = 1 - betas[0] ** t, 1 - betas[1] ** t
    p = p - (lr / bc1) * m / (v.sqrt() / bc2 ** 0.5 + eps)
    return p, m, v


@pytest.mark.gpu
def test_flat_adam_matches_reference_math():
    import ctypes
    from deepvoice3_pytorch_b200._lib import lib
    n = 100003
    torch.manual_seed(0)
    p = torch.randn(n, device="cuda")
    g = torch.randn(n, device="cuda") * 2
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    pr, mr, vr = p.clone(), m.clone(), v.clone()
    hyper = torch.zeros(4, device="cuda")
    sumsq = torch.zeros(1, device="cuda")
    scratch = torch.zeros(lib.raw("dv3_sumsq_scratch_floats")(), device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    for t in range(1, 4):
        hyper.copy_(torch.tensor([1e-3, 1 - 0.5 ** t, 1 - 0.9 ** t, 0.5]))
        lib.call("dv3_sumsq", vp(g), n, vp(sumsq), vp(scratch), st)
        assert float(sumsq) == float(sumsq.clone())
        again = torch.zeros(1, device="cuda")
        lib.call("dv3_sumsq", vp(g), n, vp(again), vp(scratch), st)
        assert float(again) == float(sumsq), "sumsq must be deterministic"
        lib.call("dv3_adam_clip", vp(p), vp(g), vp(m), vp(v), n, vp(hyper), vp(sumsq), 0.5, 0.9, 1e-6, 0.1, st)
        pr, mr, vr = adam_clip_reference(pr, g, mr, vr, t, 1e-3, (0.5, 0.9), 1e-6, 0.1, grad_scale=0.5)
    torch.testing.assert_close(p, pr, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(m, mr, rtol=1e-5, atol=1e-7)


@pytest.mark.gpu
def test_training_loss_matches_oracle_and_step_runs():
    """Device-side losses == the oracle's restatement of train.py:665-740; a few eager and graph-replayed steps
    run, stay finite and actually move the parameters."""
    from deepvoice3_pytorch_b200 import builder, ops
    from deepvoice3_pytorch_b200.train_step import (TrainStep, make_synthetic_batch, to_device, training_loss,
                                                    guided_attention_mask)
    from oracle import dv3_oracle as O
    kw = dict(n_vocab=149, embed_dim=64, mel_dim=80, linear_dim=129, r=1, downsample_step=4, kernel_size=3,
              encoder_channels=128, decoder_channels=128, converter_channels=128, use_memory_mask=True,
              key_projection=True, value_projection=True, dropout=0.05, max_positions=256)
    host = make_synthetic_batch(B=3, T_text=20, T_mel=64, linear_dim=129, seed=5)
    host["target_lengths"] = torch.tensor([64, 48, 32])
    host["input_lengths_dev"] = torch.tensor([20, 17, 9])
    host["input_lengths"] = np.array([20, 17, 9])
    for b, n in enumerate([20, 17, 9]):
        host["x"][b, n:] = 0
        host["text_positions"][b, n:] = 0
    # losses on random "outputs"
    gen = torch.Generator().manual_seed(0)
    outs = (torch.rand(3, 16, 80, generator=gen), torch.rand(3, 64, 129, generator=gen),
            torch.softmax(torch.randn(2, 3, 16, 20, generator=gen), -1), torch.rand(3, 16, 1, generator=gen))
    want = O.training_loss(outs, host["mel"], host["y"], host["done"], host["input_lengths"],
                           host["target_lengths"].numpy())
    dev = to_device(host, "cuda")
    got = training_loss(tuple(o.cuda() for o in outs), dev)
    torch.testing.assert_close(got.cpu(), want, rtol=1e-5, atol=1e-6)
    # fused loss kernels: same value, same gradients as the torch-op restatement
    from deepvoice3_pytorch_b200.train_step import fused_training_loss
    leaves = [o.cuda().requires_grad_(True) for o in outs]
    training_loss(tuple(leaves), dev).backward()
    want_g = [t.grad.clone() for t in leaves]
    leaves2 = [o.cuda().requires_grad_(True) for o in outs]
    got2 = fused_training_loss(tuple(leaves2), dev)
    torch.testing.assert_close(got2.cpu(), want, rtol=1e-5, atol=1e-6)
    got2.backward()
    for a_, b_ in zip(leaves2, want_g):
        torch.testing.assert_close(a_.grad, b_, rtol=2e-4, atol=1e-9)
    W = guided_attention_mask(dev["input_lengths_dev"], dev["target_lengths"] // 4, 16, 20, 0.2)
    np.testing.assert_allclose(W.cpu().numpy(), O.guided_attentions(host["input_lengths"], np.array([16, 12, 8]),
                                                                   16, 20, 0.2), rtol=1e-6, atol=1e-7)
    for use_graph in (False, True):
        torch.manual_seed(0)
        model = builder.deepvoice3(**kw).cuda()
        step = TrainStep(model, use_graph=use_graph)
        before = step.arena.flat.clone()
        losses = [float(step.step(dev).item()) for _ in range(4)]
        assert all(np.isfinite(losses)), losses
        assert float((step.arena.flat - before).abs().max()) > 0
        assert float(step.opt.grad_norm().item()) > 0
        ops.check_index_errors()


@pytest.mark.gpu
def test_gradient_sink_equals_autograd():
    """TrainStep lets the kernels accumulate parameter gradients straight into the flat arena (ops.grad_sink) instead
    of returning them to autograd: both routes must give the same gradients (tensor-core and exact-fp32 modes)."""
    from deepvoice3_pytorch_b200 import builder, ops
    from deepvoice3_pytorch_b200.train_step import ParameterArena, make_synthetic_batch, to_device, fused_training_loss
    kw = dict(n_vocab=149, embed_dim=64, mel_dim=80, linear_dim=129, r=1, downsample_step=4, kernel_size=3,
              encoder_channels=128, decoder_channels=128, converter_channels=128, use_memory_mask=True,
              key_projection=True, value_projection=True, dropout=0.0, max_positions=256)
    dev = to_device(make_synthetic_batch(B=2, T_text=24, T_mel=64, linear_dim=129, seed=9), "cuda")
    old_math = ops.conv_math
    try:
        for math in ("tc", "fp32"):
            ops.conv_math = math
            torch.manual_seed(0)
            model = builder.deepvoice3(**kw).cuda().train()
            arena = ParameterArena(model)

            def run(sink):
                arena.zero_grad()
                ops.grad_sink = sink
                try:
                    outs = model(dev["x"], dev["mel"], text_positions=dev["text_positions"],
                                 frame_positions=dev["frame_positions"], input_lengths=dev["input_lengths_dev"])
                    fused_training_loss(outs, dev).backward()
                finally:
                    ops.grad_sink = False
                torch.cuda.synchronize()
                return arena.grad.clone()

            g_autograd, g_sink = run(False), run(True)
            assert float(g_autograd.abs().max()) > 0
            torch.testing.assert_close(g_sink, g_autograd, rtol=1e-5, atol=1e-7)
    finally:
        ops.conv_math = old_math


@pytest.mark.gpu
def test_weight_bank_equals_per_layer_weight_norm():
    """The batched weight norm of TrainStep (WeightBank: all layers normalised / packed in two launches, all weight-norm
    backwards in one) does the same arithmetic as the per-layer kernels: same gradients (up to the atomics noise of
    the embedding / bias reductions), same losses over optimizer steps in eager and CUDA-graph mode, and the bank
    really is in use from the second pass on."""
    from deepvoice3_pytorch_b200 import builder, ops
    from deepvoice3_pytorch_b200.train_step import TrainStep, make_synthetic_batch, to_device
    kw = dict(n_vocab=149, embed_dim=64, mel_dim=80, linear_dim=129, r=1, downsample_step=4, kernel_size=3,
              encoder_channels=128, decoder_channels=128, converter_channels=128, use_memory_mask=True,
              key_projection=True, value_projection=True, dropout=0.05, max_positions=256)
    dev = to_device(make_synthetic_batch(B=2, T_text=24, T_mel=64, linear_dim=129, seed=9), "cuda")
    old_math = ops.conv_math
    ops.conv_math = "tc"
    try:
        def make(bank, graph):
            torch.manual_seed(0)
            ops.rng.manual_seed(77, torch.device("cuda"))
            return TrainStep(builder.deepvoice3(**kw).cuda().train(), use_graph=graph, weight_bank=bank)

        def grads(bank):
            step = make(bank, False)
            step._forward_backward(dev)              # first pass registers the layers (per-layer kernels)
            loss = step._forward_backward(dev)       # second pass: prepared planes + deferred weight-norm backward
            torch.cuda.synchronize()
            return step, float(loss), step.arena.grad.clone()

        _, l_ref, g_ref = grads(False)
        step, l_bank, g_bank = grads(True)
        assert len(step.bank.layers) > 10 and step.bank._fwd is not None and step.bank._bwd is not None
        assert float(g_ref.abs().max()) > 0
        assert abs(l_bank - l_ref) <= 1e-6 * abs(l_ref)
        torch.testing.assert_close(g_bank, g_ref, rtol=1e-5, atol=1e-7)

        def losses(bank, graph):
            step = make(bank, graph)
            out = [float(step.step(dev)) for _ in range(4)]
            torch.cuda.synchronize()
            return out

        ref = losses(False, False)
        for graph in (False, True):
            np.testing.assert_allclose(losses(True, graph), ref, rtol=2e-5)
    finally:
        ops.conv_math = old_math


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["step0", "step1", "step2"])
def test_fused_loss_kernels_match_reference_train_loop(name):
    """csrc/loss.cu against the loss of the reference's own train() (tests/golden/train_fns.npz ``step*``: total loss
    and d(loss)/d(every model output), incl. the priority-bin branch, w = 0 / bw = 0 and guided attention off)."""
    import golden_util as G
    from test_train_golden import step_case_inputs
    from deepvoice3_pytorch_b200.train_step import fused_training_loss
    case = G.load("train_fns.npz")[name]
    outs, batch, kw = step_case_inputs(case, "cuda")
    loss = fused_training_loss(outs, batch, **kw)
    loss.backward()
    np.testing.assert_allclose(float(loss), float(case["out"]["loss"]), rtol=1e-5)
    for o, k in zip(outs, ("mel_out", "lin_out", "attn", "done_hat")):
        np.testing.assert_allclose(o.grad.cpu().numpy(), case["out"]["grad_" + k], rtol=1e-3, atol=1e-8, err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["specloss0", "specloss1", "specloss2", "specloss3", "specloss4"])
def test_spec_loss_kernel_matches_reference_spec_loss(name):
    """dv3_spec_loss (r = 0: no frame shift) against the reference's spec_loss on the same tensors."""
    import ctypes
    import golden_util as G
    from deepvoice3_pytorch_b200._lib import lib
    case = G.load("train_fns.npz")[name]
    w, bw, pbin, pw = [float(v) for v in case["meta"]["cfg"]]
    y_hat = torch.from_numpy(case["in"]["y_hat"]).cuda()
    y = torch.from_numpy(case["in"]["y"]).cuda()
    lens = torch.from_numpy(case["in"]["lengths"]).cuda()
    B, T, D = y_hat.shape
    # the kernel pairs y_hat[:, t] with y[:, t+r] for t < T-r; emulate the un-shifted call by padding one frame
    r = 1
    y_hat_p = torch.cat([y_hat, torch.full((B, 1, D), 0.5, device="cuda")], 1).contiguous()
    y_p = torch.cat([torch.zeros(B, 1, D, device="cuda"), y], 1).contiguous()
    grad = torch.empty_like(y_hat_p)
    loss = torch.zeros(1, device="cuda")
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib.call("dv3_spec_loss", vp(y_hat_p), vp(y_p), vp((lens + r).contiguous()), vp(grad), vp(loss), B, T + 1, D, r,
             w, bw, max(int(pbin), 0), pw, st)
    want = (1 - bw) * float(case["out"]["l1"]) + bw * float(case["out"]["bd"])
    np.testing.assert_allclose(float(loss), want, rtol=1e-5)
    np.testing.assert_allclose(grad[:, :T].cpu().numpy(), case["out"]["grad"], rtol=1e-3, atol=1e-9)


@pytest.mark.gpu
def test_dropout_masks_change_every_training_forward_without_trainstep():
    """A plain model(...) / loss.backward() / optimizer.step() loop (reference train.py on this package) must see a
    fresh dropout mask every step: the model's own forward draws a new seed (ops.DropoutState.begin_forward); eval
    forwards are deterministic; the backward uses the mask of ITS forward even if another forward ran in between."""
    from deepvoice3_pytorch_b200 import builder
    kw = dict(n_vocab=149, embed_dim=64, mel_dim=80, linear_dim=129, r=1, downsample_step=4, kernel_size=3,
              encoder_channels=128, decoder_channels=128, converter_channels=128, use_memory_mask=True,
              key_projection=True, value_projection=True, dropout=0.2, max_positions=256)
    torch.manual_seed(0)
    model = builder.deepvoice3(**kw).cuda()
    host = __import__("deepvoice3_pytorch_b200.train_step", fromlist=["x"]).make_synthetic_batch(
        B=4, T_text=48, T_mel=512, linear_dim=129, seed=2)
    b = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in host.items()}

    def fwd():
        return model(b["x"], b["mel"], text_positions=b["text_positions"], frame_positions=b["frame_positions"],
                     input_lengths=b["input_lengths"])
    model.train()
    o1 = fwd()
    o2 = fwd()
    assert not torch.equal(o1[1], o2[1]), "two training forwards used the same dropout masks"
    # backward of the FIRST forward after a second forward ran: gradients must equal those of an isolated run
    g_after = torch.autograd.grad(o1[1].sum(), model.postnet.convolutions[0].weight_v, retain_graph=False)[0]
    assert torch.isfinite(g_after).all() and float(g_after.abs().max()) > 0
    model.eval()
    with torch.no_grad():
        e1, e2 = fwd(), fwd()
    assert torch.equal(e1[1], e2[1])
    # model.postnet called on its own (reference train.py:700) also draws fresh masks
    model.train()
    x = torch.rand(4, 128, model.postnet.in_dim, device="cuda")
    assert not torch.equal(model.postnet(x), model.postnet(x))


@pytest.mark.gpu
def test_train_step_checkpoint_resume():
    """TrainStep.state_dict() -> a new TrainStep.load_state_dict() continues exactly: Adam moments, bias-correction
    step and the learning-rate schedule position are restored (reference checkpoint keys, train.py:787-810)."""
    from deepvoice3_pytorch_b200 import builder, ops
    from deepvoice3_pytorch_b200.train_step import TrainStep, make_synthetic_batch, to_device
    kw = dict(n_vocab=149, embed_dim=64, mel_dim=80, linear_dim=129, r=1, downsample_step=4, kernel_size=3,
              encoder_channels=128, decoder_channels=128, converter_channels=128, use_memory_mask=True,
              key_projection=True, value_projection=True, dropout=0.0, max_positions=256)
    dev = to_device(make_synthetic_batch(B=2, T_text=24, T_mel=64, linear_dim=129, seed=9), "cuda")
    torch.manual_seed(0)
    a = TrainStep(builder.deepvoice3(**kw).cuda().train())
    for _ in range(3):
        a.step(dev)
    ckpt = a.state_dict(global_epoch=7)
    assert set(ckpt) == {"state_dict", "optimizer", "global_step", "global_epoch"} and ckpt["global_step"] == 3
    ckpt = {k: (v if not isinstance(v, dict) else __import__("copy").deepcopy(v)) for k, v in ckpt.items()}
    want = [float(a.step(dev)) for _ in range(2)]
    torch.manual_seed(123)                                   # different init: everything must come from the checkpoint
    b = TrainStep(builder.deepvoice3(**kw).cuda().train())
    assert b.load_state_dict(ckpt) == 7 and b.global_step == 3 and b.opt.t == 3
    got = [float(b.step(dev)) for _ in range(2)]
    np.testing.assert_allclose(got, want, rtol=1e-5)
    # ... and torch.optim.Adam accepts the optimizer part (a reference-side load_checkpoint)
    opt = torch.optim.Adam(b.model.get_trainable_parameters(), lr=1.0)
    opt.load_state_dict(b.opt.state_dict())
