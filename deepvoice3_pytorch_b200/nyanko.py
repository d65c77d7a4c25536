""""Nyanko" (Tachibana et al. 2017) variant with the reference's classes and signatures
(reference deepvoice3_pytorch/nyanko.py): HighwayConv1d stacks, a single attention layer, Q||R concat."""
import torch
from torch import nn

from . import ops
from .modules import (Embedding, Linear, Conv1d, ConvTranspose1d, HighwayConv1d, get_mask_from_lengths,
                      position_encoding_init, run_conv_stack)
from .deepvoice3 import AttentionLayer


def _highways(channels, kernel_size, dilations, causal, dropout):
    return [HighwayConv1d(channels, channels, kernel_size=kernel_size, padding=None, dilation=d, causal=causal,
                          std_mul=1.0, dropout=dropout) for d in dilations]


class Encoder(nn.Module):
    def __init__(self, n_vocab, embed_dim, channels, kernel_size=3, n_speakers=1, speaker_embed_dim=16,
                 embedding_weight_std=0.01, padding_idx=None, dropout=0.1):
        super().__init__()
        self.dropout = dropout
        self.embed_tokens = Embedding(n_vocab, embed_dim, padding_idx, embedding_weight_std)
        E, D2 = embed_dim, 2 * channels
        self.convnet = nn.Sequential(                                   # reference nyanko.py:29-58
            Conv1d(E, D2, kernel_size=1, padding=0, dilation=1, std_mul=1.0),
            nn.ReLU(inplace=True),
            Conv1d(D2, D2, kernel_size=1, padding=0, dilation=1, std_mul=2.0),
            *_highways(D2, kernel_size, [1, 3, 9, 27, 1, 3, 9, 27, 1, 1], False, dropout),
            HighwayConv1d(D2, D2, kernel_size=1, padding=0, dilation=1, std_mul=1.0, dropout=dropout),
        )

    def forward(self, text_sequences, text_positions=None, lengths=None, speaker_embed=None):
        x = self.embed_tokens(text_sequences)
        x = ops.transpose12(run_conv_stack(self.convnet, ops.transpose12(x)))
        keys, values = x.split(x.size(-1) // 2, dim=-1)
        return keys, values


class Decoder(nn.Module):
    def __init__(self, embed_dim, in_dim=80, r=5, channels=256, kernel_size=3, n_speakers=1,
                 speaker_embed_dim=16, max_positions=512, padding_idx=None, dropout=0.1, use_memory_mask=False,
                 force_monotonic_attention=False, query_position_rate=1.0, key_position_rate=1.29,
                 window_ahead=3, window_backward=1, key_projection=False, value_projection=False):
        super().__init__()
        self.dropout = dropout
        self.in_dim = in_dim
        self.r = r
        D, Fr = channels, in_dim * r

        def c1(cin, cout, std_mul):
            return Conv1d(cin, cout, kernel_size=1, padding=0, dilation=1, std_mul=std_mul)

        self.audio_encoder_modules = nn.ModuleList(                     # reference nyanko.py:95-121
            [c1(Fr, D, 1.0), nn.ReLU(inplace=True), c1(D, D, 2.0), nn.ReLU(inplace=True), c1(D, D, 2.0)]
            + _highways(D, kernel_size, [1, 3, 9, 27, 1, 3, 9, 27, 3, 3], True, dropout))
        self.attention = AttentionLayer(D, D, dropout=dropout, window_ahead=window_ahead,
                                        window_backward=window_backward, key_projection=key_projection,
                                        value_projection=value_projection)
        self.audio_decoder_modules = nn.ModuleList(                     # reference nyanko.py:129-152
            [c1(2 * D, D, 1.0)] + _highways(D, kernel_size, [1, 3, 9, 27, 1, 1], True, dropout)
            + [c1(D, D, 1.0), nn.ReLU(inplace=True), c1(D, D, 2.0), nn.ReLU(inplace=True), c1(D, D, 2.0),
               nn.ReLU(inplace=True)])
        self.last_conv = c1(D, Fr, 2.0)
        self.fc = Linear(Fr, 1)

        # position rates are baked into plain embeddings here (reference nyanko.py:162-169)
        self.embed_query_positions = Embedding(max_positions, D, padding_idx)
        self.embed_query_positions.weight.data = position_encoding_init(
            max_positions, D, position_rate=query_position_rate, sinusoidal=True)
        self.embed_keys_positions = Embedding(max_positions, D, padding_idx)
        self.embed_keys_positions.weight.data = position_encoding_init(
            max_positions, D, position_rate=key_position_rate, sinusoidal=True)

        self.max_decoder_steps = 200
        self.min_decoder_steps = 10
        self.use_memory_mask = use_memory_mask
        self.force_monotonic_attention = force_monotonic_attention

    def forward(self, encoder_out, inputs=None, text_positions=None, frame_positions=None, speaker_embed=None,
                lengths=None):
        """Teacher-forced decoder (reference nyanko.py:177-248)."""
        if inputs is None:                 # inference (reference nyanko.py:180-184)
            assert text_positions is not None
            self.start_fresh_sequence()
            return self.incremental_forward(encoder_out, text_positions)
        if inputs.size(-1) == self.in_dim:
            inputs = inputs.reshape(inputs.size(0), inputs.size(1) // self.r, -1)
        assert inputs.size(-1) == self.in_dim * self.r
        keys, values = encoder_out
        mask = get_mask_from_lengths(keys, lengths) if (self.use_memory_mask and lengths is not None) else None
        if text_positions is not None:
            keys = keys + self.embed_keys_positions(text_positions)
        keys_bct = ops.transpose12(keys)
        values_bct = ops.transpose12(values)

        x = run_conv_stack(self.audio_encoder_modules, ops.transpose12(inputs))
        Q = x
        q = x if frame_positions is None else x + ops.transpose12(self.embed_query_positions(frame_positions))
        R, alignments = self.attention.forward_bct(q, keys_bct, values_bct, mask)
        x = run_conv_stack(self.audio_decoder_modules, torch.cat((R, Q), dim=1))
        decoder_states = ops.transpose12(x)
        x = ops.transpose12(self.last_conv(x))
        outputs = torch.sigmoid(x)
        done = torch.sigmoid(self.fc(x))
        return outputs, alignments.unsqueeze(0), done, decoder_states

    def incremental_forward(self, encoder_out, text_positions, initial_input=None, test_inputs=None):
        """Autoregressive decoding (reference nyanko.py:250-338); see incremental.py."""
        from .incremental import decode
        return decode(self, encoder_out, text_positions, None, initial_input, test_inputs)

    def start_fresh_sequence(self):
        for m in list(self.audio_encoder_modules) + list(self.audio_decoder_modules) + [self.last_conv]:
            if hasattr(m, "clear_buffer"):
                m.clear_buffer()


class Converter(nn.Module):
    def __init__(self, in_dim, out_dim, channels=512, kernel_size=3, dropout=0.1):
        super().__init__()
        self.dropout = dropout
        self.in_dim = in_dim
        self.out_dim = out_dim
        C, Fd = channels, out_dim

        def c1(cin, cout, std_mul=1.0):
            return Conv1d(cin, cout, kernel_size=1, padding=0, dilation=1, std_mul=std_mul)

        def up():
            return ConvTranspose1d(C, C, kernel_size=2, padding=0, stride=2, std_mul=1.0)

        self.convnet = nn.Sequential(                                   # reference nyanko.py:363-399
            c1(in_dim, C), *_highways(C, kernel_size, [1, 3], False, dropout),
            up(), *_highways(C, kernel_size, [1, 3], False, dropout),
            up(), *_highways(C, kernel_size, [1, 3], False, dropout),
            c1(C, 2 * C), *_highways(2 * C, kernel_size, [1, 1], False, dropout),
            c1(2 * C, Fd),
            c1(Fd, Fd), nn.ReLU(inplace=True), c1(Fd, Fd, 2.0), nn.ReLU(inplace=True),
            c1(Fd, Fd, 2.0), nn.Sigmoid(),
        )

    @ops.forward_scope
    def forward(self, x, speaker_embed=None):
        return ops.transpose12(run_conv_stack(self.convnet, ops.transpose12(x)))
