"""GPU parity of the inference path: the device-resident step program of
deepvoice3_pytorch_b200/incremental.py against the live reference's Decoder.incremental_forward (golden fixtures
tests/golden/incremental.npz), against the teacher-forced batch decoder (the reference's own
test_incremental_correctness, tests/test_deepvoice3.py:184-235, atol 1e-5) and, module by module, the reference's
tests/test_conv.py (incremental == batch convolution, exactly)."""
import numpy as np
import pytest
import torch

import golden_util as G
from test_gpu_models import _build, preset_kwargs

pytestmark = pytest.mark.gpu
INCR = G.load("incremental.npz")
RTOL, ATOL = 1e-4, 1e-5


def close(a, b, what, rtol=RTOL, atol=ATOL):
    a = a.detach().float().cpu().numpy() if torch.is_tensor(a) else a
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=what)


def _model(case):
    model = _build(G.kwargs_of(case)).cuda()
    model.load_state_dict(G.tensors(case["sd"]), strict=True)
    model.eval()
    dec = model.seq2seq.decoder
    dec.max_decoder_steps = int(case["meta"]["max_decoder_steps"])
    dec.min_decoder_steps = int(case["meta"]["min_decoder_steps"])
    return model


def _decode(model, case, mode, **kw):
    ins = G.tensors(case["in"], "cuda")
    dec = model.seq2seq.decoder
    enc = (ins["keys"], ins["values"])
    test_inputs = ins["mel"] if mode == "forced" else None
    if hasattr(dec, "audio_encoder_modules"):
        return dec.incremental_forward(enc, ins["text_positions"], test_inputs=test_inputs, **kw)
    spk = model.embed_speakers(ins["speaker_ids"]) if "speaker_ids" in ins else None
    return dec.incremental_forward(enc, ins["text_positions"], spk, test_inputs=test_inputs, **kw)


@pytest.mark.parametrize("mode", ["forced", "free"])
@pytest.mark.parametrize("name", list(INCR))
def test_incremental_decoder_golden(name, mode):
    """Same number of steps and same values as the reference decoder, teacher-forced and free-running (monotonic
    window on/off, single/multi-speaker, deepvoice3 and nyanko)."""
    case = INCR[name]
    model = _model(case)
    outputs, alignments, dones, states = _decode(model, case, mode)
    ref = case[mode]
    assert len(dones) == ref["dones"].shape[1] and tuple(dones[0].shape) == (outputs.size(0), 1, 1)
    close(outputs, ref["outputs"], "outputs")
    close(alignments, ref["alignments"], "alignments")
    close(torch.cat(dones, dim=1), ref["dones"], "dones")
    close(states, ref["states"], "decoder_states")


@pytest.mark.parametrize("name", list(INCR))
def test_incremental_matches_oracle_and_graph_matches_eager(name):
    """The CPU oracle's stepwise decoder on the same inputs, and CUDA-graph replay == eager stepping bit for bit."""
    from test_oracle_golden import _incremental_oracle
    case = INCR[name]
    model = _model(case)
    for mode in ("forced", "free"):
        want = _incremental_oracle(case, mode)
        got_g = _decode(model, case, mode)
        for g, w, key in zip((got_g[0], got_g[1], torch.cat(got_g[2], dim=1), got_g[3]), want,
                             ("outputs", "alignments", "dones", "states")):
            close(g, w.numpy(), "%s %s" % (mode, key))
    from deepvoice3_pytorch_b200 import incremental
    ins = G.tensors(case["in"], "cuda")
    dec = model.seq2seq.decoder
    spk = model.embed_speakers(ins["speaker_ids"]) if "speaker_ids" in ins else None
    a = incremental.decode(dec, (ins["keys"], ins["values"]), ins["text_positions"], spk, use_graph=True)
    b = incremental.decode(dec, (ins["keys"], ins["values"]), ins["text_positions"], spk, use_graph=False)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[3], b[3]) and len(a[2]) == len(b[2])


@pytest.mark.parametrize("name", list(INCR))
def test_model_inference_call_golden(name):
    """The user-facing inference call model(text, text_positions=..., speaker_ids=...) (reference synthesis.py:62-64):
    encoder -> free-running decoder -> converter.  Same step count; values within the north_star tolerance (the
    feedback loop runs through our own encoder and conv kernels, so differences of a few 1e-6 per step add up)."""
    from deepvoice3_pytorch_b200 import ops
    case = INCR[name]
    model = _model(case)
    ins = G.tensors(case["in"], "cuda")
    old = ops.conv_math
    ops.conv_math = "fp32"
    try:
        with torch.no_grad():
            mel, linear, alignments, done = model(ins["text"], text_positions=ins["text_positions"],
                                                  speaker_ids=ins.get("speaker_ids"))
    finally:
        ops.conv_math = old
    ref = case["model"]
    assert tuple(mel.shape) == tuple(ref["mel"].shape)
    close(mel, ref["mel"], "mel", rtol=1e-3, atol=1e-4)
    close(linear, ref["linear"], "linear", rtol=1e-3, atol=1e-4)
    close(alignments, ref["alignments"], "alignments", rtol=1e-3, atol=1e-4)
    close(torch.cat(done, dim=1), ref["dones"], "dones", rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("preset", ["deepvoice3_ljspeech", "nyanko_ljspeech"])
def test_teacher_forced_incremental_equals_batch_decoder(preset):
    """reference tests/test_deepvoice3.py:184-235 / tests/test_nyanko.py:86-133 at the PRESET sizes: feeding the
    target frames one at a time through the step program reproduces the teacher-forced batch decoder (atol 1e-5 in
    the reference's test)."""
    from deepvoice3_pytorch_b200 import builder, ops
    bname, kw = preset_kwargs(preset)
    kw = dict(kw, force_monotonic_attention=False, use_memory_mask=False)
    torch.manual_seed(5)
    model = getattr(builder, bname)(dropout=0.0, **kw).cuda().eval()
    B, Tt, Td = 2, 40, 48
    gen = torch.Generator().manual_seed(11)
    text = torch.randint(2, 149, (B, Tt), generator=gen).cuda()
    tpos = torch.arange(1, Tt + 1)[None].repeat(B, 1).cuda()
    fpos = torch.arange(1, Td + 1)[None].repeat(B, 1).cuda()
    mel = torch.rand(B, Td, 80, generator=gen).cuda()
    old = ops.conv_math
    ops.conv_math = "fp32"
    try:
        with torch.no_grad():
            enc = model.seq2seq.encoder(text)
            dec = model.seq2seq.decoder
            want = dec(enc, mel, text_positions=tpos, frame_positions=fpos)
            dec.start_fresh_sequence()
            got = dec.incremental_forward(enc, tpos, test_inputs=mel)
    finally:
        ops.conv_math = old
    close(got[0], want[0].cpu().numpy(), "outputs", atol=2e-5)
    close(torch.cat(got[2], dim=1), want[2].cpu().numpy(), "done", atol=2e-5)
    close(got[3], want[3].cpu().numpy(), "decoder_states", atol=5e-5)
    n_att = want[1].size(0)
    close(got[1], (want[1][0] * (2 ** (n_att - 1)) / n_att).cpu().numpy(), "alignment (first layer, scaled)",
          atol=2e-5)


def test_module_level_incremental_forward_equals_batch_conv():
    """reference tests/test_conv.py:10-63: Conv1d.incremental_forward frame by frame == the causal batch convolution,
    exactly (ramp input, unit weights), for the reference's (kernel_size, dilation) grid; plus the gated blocks."""
    from deepvoice3_pytorch_b200 import modules
    for B, T, C in [(1, 10, 2), (4, 10, 4)]:
        for k in [2, 3]:
            for d in [1, 2, 3, 4, 5, 9, 27]:
                conv = modules.Conv1d(C, 2 * C, k, dilation=d, padding=(k - 1) * d).cuda().eval()
                with torch.no_grad():
                    conv.weight_v.fill_(1.0)
                    conv.weight_g.copy_(torch.norm_except_dim(conv.weight_v, 2, 0))
                    conv.bias.zero_()
                x = (torch.zeros(B, C, T) + torch.arange(0, T).float()).cuda()
                with torch.no_grad():
                    y = conv(x, causal=True)
                    conv.clear_buffer()
                    steps = [conv.incremental_forward(x[:, :, t:t + 1].transpose(1, 2).contiguous())
                             for t in range(T)]
                y_inc = torch.cat(steps, dim=1).transpose(1, 2)
                torch.testing.assert_close(y_inc, y, rtol=1e-6, atol=1e-5)
    torch.manual_seed(3)
    for blk in (modules.Conv1dGLU(1, 16, 32, 32, 3, dropout=0.0, dilation=3, causal=True, residual=True),
                modules.HighwayConv1d(32, 32, kernel_size=3, dilation=9, causal=True, dropout=0.0)):
        blk = blk.cuda().eval()
        x = torch.randn(2, 32, 20, device="cuda")
        with torch.no_grad():
            y = blk(x)
            blk.clear_buffer()
            steps = [blk.incremental_forward(x[:, :, t:t + 1].transpose(1, 2).contiguous()) for t in range(20)]
        torch.testing.assert_close(torch.cat(steps, dim=1).transpose(1, 2), y, rtol=1e-5, atol=1e-5)
