"""deepvoice3_pytorch_b200 -- B200-native drop-in for the training hot path of r9y9/deepvoice3_pytorch.

Mirrors the reference package surface (reference deepvoice3_pytorch/__init__.py): ``MultiSpeakerTTSModel``,
``AttentionSeq2Seq`` and ``builder.{deepvoice3, nyanko, deepvoice3_multispeaker}``, with identical
``state_dict`` keys.  All arithmetic runs in hand-written sm_100a kernels behind a C ABI
(include/dv3b200.h, csrc/); there is no CPU fallback.
"""
__version__ = "0.1.0"

from torch import nn

from . import ops
from .modules import Embedding


class MultiSpeakerTTSModel(nn.Module):
    """Attention seq2seq model + post processing network (reference __init__.py:11-97)."""

    def __init__(self, seq2seq, postnet, mel_dim=80, linear_dim=513, n_speakers=1, speaker_embed_dim=16,
                 padding_idx=None, trainable_positional_encodings=False,
                 use_decoder_state_for_postnet_input=False, speaker_embedding_weight_std=0.01,
                 freeze_embedding=False):
        super().__init__()
        self.seq2seq = seq2seq
        self.postnet = postnet
        self.mel_dim = mel_dim
        self.linear_dim = linear_dim
        self.trainable_positional_encodings = trainable_positional_encodings
        self.use_decoder_state_for_postnet_input = use_decoder_state_for_postnet_input
        self.freeze_embedding = freeze_embedding
        if n_speakers > 1:
            self.embed_speakers = Embedding(n_speakers, speaker_embed_dim, padding_idx=None,
                                            std=speaker_embedding_weight_std)
        self.n_speakers = n_speakers
        self.speaker_embed_dim = speaker_embed_dim

    def make_generation_fast_(self):
        raise NotImplementedError("weight-norm folding for inference is outside the training hot path")

    def get_trainable_parameters(self):
        """Everything except the position tables (unless trainable) and, optionally, the text embedding
        (reference __init__.py:48-63)."""
        frozen = set()
        encoder, decoder = self.seq2seq.encoder, self.seq2seq.decoder
        if not self.trainable_positional_encodings:
            frozen |= set(map(id, decoder.embed_query_positions.parameters()))
            frozen |= set(map(id, decoder.embed_keys_positions.parameters()))
        if self.freeze_embedding:
            frozen |= set(map(id, encoder.embed_tokens.parameters()))
        return (p for p in self.parameters() if id(p) not in frozen)

    def forward(self, text_sequences, mel_targets=None, speaker_ids=None, text_positions=None,
                frame_positions=None, input_lengths=None):
        """-> mel_outputs (B,T,mel_dim), linear_outputs (B,T*ds,linear_dim), alignments (N,B,T_dec,T_text),
        done (B,T_dec,1)."""
        ops.rng.start_forward()
        B = text_sequences.size(0)
        if speaker_ids is not None:
            assert self.n_speakers > 1
            speaker_embed = self.embed_speakers(speaker_ids)
        else:
            speaker_embed = None
        mel_outputs, alignments, done, decoder_states = self.seq2seq(
            text_sequences, mel_targets, speaker_embed, text_positions, frame_positions, input_lengths)
        mel_outputs = mel_outputs.reshape(B, -1, self.mel_dim)
        if self.use_decoder_state_for_postnet_input:
            postnet_inputs = decoder_states.reshape(B, mel_outputs.size(1), -1)
        else:
            postnet_inputs = mel_outputs
        linear_outputs = self.postnet(postnet_inputs, speaker_embed)
        assert linear_outputs.size(-1) == self.linear_dim
        return mel_outputs, linear_outputs, alignments, done


class AttentionSeq2Seq(nn.Module):
    """Encoder + Decoder with attention (reference __init__.py:100-126)."""

    def __init__(self, encoder, decoder):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        if isinstance(self.decoder.attention, nn.ModuleList):
            self.encoder.num_attention_layers = sum(layer is not None for layer in decoder.attention)

    def forward(self, text_sequences, mel_targets=None, speaker_embed=None, text_positions=None,
                frame_positions=None, input_lengths=None):
        encoder_outputs = self.encoder(text_sequences, lengths=input_lengths, speaker_embed=speaker_embed)
        return self.decoder(encoder_outputs, mel_targets, text_positions=text_positions,
                            frame_positions=frame_positions, speaker_embed=speaker_embed, lengths=input_lengths)
