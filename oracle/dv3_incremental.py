"""TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's INFERENCE path: autoregressive decoding one frame
at a time (reference conv.py:17-46 ring buffer x linearised weight, modules.py:142-167 / 197-226 gate epilogues on a
(B,1,C) slice, deepvoice3.py:132-176 attention with the monotonic window, decoder loops deepvoice3.py:367-485 and
nyanko.py:250-338) -- a pure function of a reference-keyed state_dict, a spec (oracle/specs.py) and the inputs.

PARITY: pinned.  ``tests/golden/make_golden.py incremental`` runs the live reference decoders (teacher-forced and
free-running, single/multi-speaker, deepvoice3 and nyanko) and stores their outputs in tests/golden/incremental.npz;
tests/test_oracle_golden.py checks every function below against them.
"""
import math

import torch
import torch.nn.functional as F

from . import dv3_oracle as O

SQRT_HALF = math.sqrt(0.5)


class IncConv:
    """reference conv.py:17-46: input_buffer of kw + (kw-1)(dilation-1) frames, shifted left every step; the frames
    at stride ``dilation`` times the (Cout, kw*Cin) linearised weight."""

    def __init__(self, sd, prefix, k=1, dilation=1):
        w = O._w(sd, prefix)
        if w.dim() == 2:
            w = w.unsqueeze(-1)
        self.k, self.d = w.size(2), dilation
        assert self.k == k
        self.w = w.transpose(1, 2).contiguous().view(w.size(0), -1)       # conv.py:55-60
        self.bias = sd[prefix + ".bias"]
        self.buffer = None

    def step(self, x):                                                     # x (B, 1, Cin)
        B = x.size(0)
        if self.k > 1:
            if self.buffer is None:
                self.buffer = x.new_zeros(B, self.k + (self.k - 1) * (self.d - 1), x.size(2))
            else:
                self.buffer[:, :-1, :] = self.buffer[:, 1:, :].clone()
            self.buffer[:, -1, :] = x[:, -1, :]
            x = self.buffer[:, 0::self.d, :].contiguous()
        return F.linear(x.reshape(B, -1), self.w, self.bias).view(B, 1, -1)


def make_stack(sd, prefix, layers, speaker_embed=None):
    """spec layer list -> list of step closures x (B,1,C) -> (B,1,C')."""
    steps = []
    for layer in layers:
        kind, idx = layer[0], layer[1]
        p = "%s.%d" % (prefix, idx)
        if kind == "conv":
            conv = IncConv(sd, p, layer[4], layer[5])
            steps.append(conv.step)
        elif kind == "relu":
            steps.append(F.relu)
        elif kind == "glu":
            _, _, C, k, d, causal, residual = layer
            conv = IncConv(sd, p + ".conv", k, d)
            soft = None
            if speaker_embed is not None and (p + ".speaker_proj.weight_v") in sd:
                soft = F.softsign(O.linear(sd, p + ".speaker_proj", speaker_embed)).unsqueeze(1)

            def glu(x, conv=conv, soft=soft, residual=residual):           # modules.py:145-164
                a, b = conv.step(x).split(x.size(-1), dim=-1)
                if soft is not None:
                    a = a + soft
                y = a * torch.sigmoid(b)
                return (y + x) * SQRT_HALF if residual else y
            steps.append(glu)
        elif kind == "hw":
            _, _, C, k, d, causal = layer
            conv = IncConv(sd, p + ".conv", k, d)

            def hw(x, conv=conv):                                          # modules.py:200-226
                a, b = conv.step(x).split(x.size(-1), dim=-1)
                T = torch.sigmoid(b)
                return T * a + (1 - T) * x
            steps.append(hw)
        else:
            raise ValueError(kind)
    return steps


def attention_step(sd, prefix, query, keys_bct, values, last_attended, window_backward=1, window_ahead=3):
    """reference deepvoice3.py:132-176 on a single query frame, with the monotonic window (150-156)."""
    residual = query
    if (prefix + ".value_projection.weight_v") in sd:
        values = O.linear(sd, prefix + ".value_projection", values)
    if (prefix + ".key_projection.weight_v") in sd:
        keys_bct = O.linear(sd, prefix + ".key_projection", keys_bct.transpose(1, 2)).transpose(1, 2)
    x = torch.bmm(O.linear(sd, prefix + ".query_projection", query), keys_bct)
    if last_attended is not None:
        backward = last_attended - window_backward
        if backward > 0:
            x[:, :, :backward] = -float("inf")
        ahead = last_attended + window_ahead
        if ahead < x.size(-1):
            x[:, :, ahead:] = -float("inf")
    probs = F.softmax(x, dim=-1)
    x = torch.bmm(probs, values)
    s = values.size(1)
    x = x * (s * math.sqrt(1.0 / s))
    x = O.linear(sd, prefix + ".out_projection", x)
    return (x + residual) * SQRT_HALF, probs


def _loop(step_fn, B, Fr, test_inputs, initial_input, min_steps, max_steps, like):
    """The while-loop both reference decoders share (deepvoice3.py:399-470, nyanko.py:274-323)."""
    outputs, alignments, dones, states = [], [], [], []
    t = 0
    cur = like.new_zeros(B, 1, Fr) if initial_input is None else initial_input
    while True:
        if test_inputs is not None:
            if t >= test_inputs.size(1):
                break
            cur = test_inputs[:, t, :].unsqueeze(1)
        elif t > 0:
            cur = outputs[-1]
        out, ali, done, state = step_fn(cur, t)
        outputs.append(out); alignments.append(ali); dones.append(done); states.append(state)
        t += 1
        if test_inputs is None:
            if (done > 0.5).all() and t > min_steps:
                break
            elif t > max_steps:
                break
    sq = lambda xs: torch.stack([x.squeeze(1) for x in xs]).transpose(0, 1).contiguous()
    return sq(outputs), sq(alignments), dones, sq(states)


@torch.no_grad()
def dv3_decoder_incremental(sd, spec, encoder_out, text_positions, speaker_embed=None, initial_input=None,
                            test_inputs=None, force_monotonic_attention=True, window_backward=1, window_ahead=3,
                            min_decoder_steps=10, max_decoder_steps=200, prefix="seq2seq.decoder"):
    """reference deepvoice3.py:367-485."""
    keys, values = encoder_out
    B = keys.size(0)
    w = spec["key_position_rate"]
    if speaker_embed is not None:
        w = w * torch.sigmoid(O.linear(sd, prefix + ".speaker_proj1", speaker_embed)).view(-1)
    keys = keys + O.sinusoidal_encoding(sd[prefix + ".embed_keys_positions.weight"], text_positions, w)
    keys = keys.transpose(1, 2).contiguous()
    pre = make_stack(sd, prefix + ".preattention", spec["preattention"], speaker_embed)
    convs = [make_stack(sd, prefix + ".convolutions", [layer], speaker_embed)[0] for layer in spec["decoder"]]
    n_att = sum(spec["attention"])
    fm = force_monotonic_attention
    if isinstance(fm, bool):
        fm = [fm] * len(spec["decoder"])
    last_attended = [0 if v else None for v in fm]
    last_conv = IncConv(sd, prefix + ".last_conv", 1, 1)

    def step(cur, t):
        frame_pos = torch.full((B, 1), t + 1, dtype=torch.long)
        wq = spec["query_position_rate"]
        if speaker_embed is not None:
            wq = wq * torch.sigmoid(O.linear(sd, prefix + ".speaker_proj2", speaker_embed)).view(-1)
        frame_pos_embed = O.sinusoidal_encoding(sd[prefix + ".embed_query_positions.weight"], frame_pos, wq)
        x = cur
        for f in pre:
            x = f(x)
        ave = None
        for idx, (f, layer, has_att) in enumerate(zip(convs, spec["decoder"], spec["attention"])):
            residual = x
            x = f(x)
            if has_att:
                x = x + frame_pos_embed
                x, ali = attention_step(sd, "%s.attention.%d" % (prefix, layer[1]), x, keys, values,
                                        last_attended[idx], window_backward, window_ahead)
                if fm[idx]:
                    last_attended[idx] = int(ali.max(-1)[1].view(-1)[0])
                ave = ali if ave is None else ave + ave                    # sic: deepvoice3.py:446
            x = (x + residual) * SQRT_HALF
        state = x
        x = last_conv.step(x)
        ave = ave / n_att
        return torch.sigmoid(x), ave, torch.sigmoid(O.linear(sd, prefix + ".fc", x)), state

    return _loop(step, B, spec["mel_dim"] * spec["r"], test_inputs, initial_input, min_decoder_steps,
                 max_decoder_steps, keys)


@torch.no_grad()
def nyanko_decoder_incremental(sd, spec, encoder_out, text_positions, initial_input=None, test_inputs=None,
                               force_monotonic_attention=True, window_backward=1, window_ahead=3,
                               min_decoder_steps=10, max_decoder_steps=200, prefix="seq2seq.decoder"):
    """reference nyanko.py:250-338."""
    keys, values = encoder_out
    B = keys.size(0)
    if text_positions is not None:
        keys = keys + F.embedding(text_positions, sd[prefix + ".embed_keys_positions.weight"],
                                  padding_idx=spec["padding_idx"])
    keys = keys.transpose(1, 2).contiguous()
    enc = make_stack(sd, prefix + ".audio_encoder_modules", spec["audio_encoder"])
    dec = make_stack(sd, prefix + ".audio_decoder_modules", spec["audio_decoder"])
    last_conv = IncConv(sd, prefix + ".last_conv", 1, 1)
    state = {"la": 0 if force_monotonic_attention else None}

    def step(cur, t):
        frame_pos = torch.full((B, 1), t + 1, dtype=torch.long)
        frame_pos_embed = F.embedding(frame_pos, sd[prefix + ".embed_query_positions.weight"],
                                      padding_idx=spec["padding_idx"])
        x = cur
        for f in enc:
            x = f(x)
        Q = x
        R, ali = attention_step(sd, prefix + ".attention", x + frame_pos_embed, keys, values, state["la"],
                                window_backward, window_ahead)
        if force_monotonic_attention:
            state["la"] = int(ali.max(-1)[1].view(-1)[0])
        x = torch.cat((R, Q), dim=-1)
        for f in dec:
            x = f(x)
        st = x
        x = last_conv.step(x)
        return torch.sigmoid(x), ali, torch.sigmoid(O.linear(sd, prefix + ".fc", x)), st

    return _loop(step, B, spec["mel_dim"] * spec["r"], test_inputs, initial_input, min_decoder_steps,
                 max_decoder_steps, keys)
