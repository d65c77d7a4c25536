"""TEST INFRASTRUCTURE ONLY.  Layer recipes of the reference builders as flat lists.

Each network is a list of layer tuples, in execution order; ``idx`` is the index of the module in
the reference's ModuleList/Sequential (that is the number that appears in the state_dict key).

    ("conv",  idx, cin, cout, k, dilation)            weight-normed Conv1d (reference modules.py:94-100)
    ("convT", idx, cin, cout)                          weight-normed ConvTranspose1d k=2,s=2 (modules.py:103-109)
    ("relu",  idx) / ("sigmoid", idx)
    ("glu",   idx, C, k, dilation, causal, residual)   Conv1dGLU (modules.py:112-167)
    ("hw",    idx, C, k, dilation, causal)             HighwayConv1d (modules.py:170-229)

Recipes restate reference deepvoice3_pytorch/builder.py:7-93 (deepvoice3), :96-169 (nyanko),
:172-258 (deepvoice3_multispeaker) plus the constructors they call
(deepvoice3.py:24-67, 179-275, 501-580; nyanko.py:15-58, 74-175, 354-399).
"""


def _dv3_stack(in_channels, convolutions, causal, residual, start_idx=0):
    """deepvoice3.py:44-61 / 214-231: 1x1+ReLU whenever the width changes, then a GLU block."""
    layers, idx = [], start_idx
    for (out_channels, k, d) in convolutions:
        if in_channels != out_channels:
            layers.append(("conv", idx, in_channels, out_channels, 1, 1)); idx += 1
            layers.append(("relu", idx)); idx += 1
            in_channels = out_channels
        layers.append(("glu", idx, out_channels, k, d, causal, residual)); idx += 1
    return layers, idx, in_channels


def deepvoice3_spec(n_vocab, embed_dim=256, mel_dim=80, linear_dim=513, r=4, downsample_step=1,
                    n_speakers=1, speaker_embed_dim=16, padding_idx=0, kernel_size=5,
                    encoder_channels=128,
                    decoder_channels=256, converter_channels=256, query_position_rate=1.0,
                    key_position_rate=1.29, use_memory_mask=False,
                    use_decoder_state_for_postnet_input=True, max_positions=512,
                    key_projection=False, value_projection=False, multispeaker_recipe=False,
                    **_unused):
    k = kernel_size
    h = encoder_channels
    dil = [1, 3, 9, 27, 1, 3, 9, 27, 1, 3]
    enc, idx, cin = _dv3_stack(embed_dim, [(h, k, d) for d in dil], causal=False, residual=True)
    enc.append(("conv", idx, cin, embed_dim, 1, 1))

    h = decoder_channels
    pre_cfg = [(h, k, 1)] if multispeaker_recipe else [(h, k, 1), (h, k, 3)]
    pre, _, cin = _dv3_stack(mel_dim * r, pre_cfg, causal=True, residual=True)
    dec = [("glu", i, h, k, d, True, False) for i, d in enumerate([1, 3, 9, 27, 1])]
    attention = [True, False, False, False, False] if multispeaker_recipe \
        else [True, False, False, False, True]

    time_upsampling = max(downsample_step // r, 1)
    in_dim = h // r if use_decoder_state_for_postnet_input else mel_dim
    c = converter_channels
    if time_upsampling == 4:
        post = [("conv", 0, in_dim, c, 1, 1), ("convT", 1, c, c),
                ("glu", 2, c, 3, 1, False, True), ("glu", 3, c, 3, 3, False, True),
                ("convT", 4, c, c),
                ("glu", 5, c, 3, 1, False, True), ("glu", 6, c, 3, 3, False, True)]
    elif time_upsampling == 2:
        post = [("conv", 0, in_dim, c, 1, 1), ("convT", 1, c, c),
                ("glu", 2, c, 3, 1, False, True), ("glu", 3, c, 3, 3, False, True)]
    elif time_upsampling == 1:
        post = [("conv", 0, in_dim, c, 1, 1), ("glu", 1, c, 3, 3, False, True)]
    else:
        raise ValueError("Not supported")
    tail, idx, cin = _dv3_stack(c, [(c, k, 1), (c, k, 3), (2 * c, k, 1), (2 * c, k, 3)],
                                causal=False, residual=True, start_idx=len(post))
    post = post + tail + [("conv", idx, cin, linear_dim, 1, 1)]

    return dict(kind="deepvoice3", n_vocab=n_vocab, embed_dim=embed_dim, mel_dim=mel_dim,
                linear_dim=linear_dim, r=r, n_speakers=n_speakers, padding_idx=padding_idx,
                speaker_embed_dim=speaker_embed_dim, encoder=enc, preattention=pre,
                decoder=dec, attention=attention, converter=post,
                decoder_channels=decoder_channels,
                query_position_rate=query_position_rate, key_position_rate=key_position_rate,
                use_memory_mask=use_memory_mask,
                use_decoder_state_for_postnet_input=use_decoder_state_for_postnet_input,
                key_projection=key_projection, value_projection=value_projection,
                max_positions=max_positions)


def nyanko_spec(n_vocab, embed_dim=128, mel_dim=80, linear_dim=513, r=1, downsample_step=4,
                n_speakers=1, padding_idx=0, kernel_size=3, encoder_channels=256,
                decoder_channels=256,
                converter_channels=512, query_position_rate=1.0, key_position_rate=1.29,
                use_memory_mask=False, use_decoder_state_for_postnet_input=False,
                max_positions=512, key_projection=False, value_projection=False, **_unused):
    assert encoder_channels == decoder_channels
    if n_speakers != 1:
        raise ValueError("Multi-speaker is not supported")
    if not (downsample_step == 4 and r == 1):
        raise ValueError("Not supported. You need to change hardcoded parameters")
    k = kernel_size
    E, D = embed_dim, encoder_channels
    enc = [("conv", 0, E, 2 * D, 1, 1), ("relu", 1), ("conv", 2, 2 * D, 2 * D, 1, 1)]
    for i, d in enumerate([1, 3, 9, 27, 1, 3, 9, 27, 1, 1]):
        enc.append(("hw", 3 + i, 2 * D, k, d, False))
    enc.append(("hw", 13, 2 * D, 1, 1, False))

    D = decoder_channels
    F = mel_dim * r
    aenc = [("conv", 0, F, D, 1, 1), ("relu", 1), ("conv", 2, D, D, 1, 1), ("relu", 3),
            ("conv", 4, D, D, 1, 1)]
    for i, d in enumerate([1, 3, 9, 27, 1, 3, 9, 27, 3, 3]):
        aenc.append(("hw", 5 + i, D, k, d, True))
    adec = [("conv", 0, 2 * D, D, 1, 1)]
    for i, d in enumerate([1, 3, 9, 27, 1, 1]):
        adec.append(("hw", 1 + i, D, k, d, True))
    adec += [("conv", 7, D, D, 1, 1), ("relu", 8), ("conv", 9, D, D, 1, 1), ("relu", 10),
             ("conv", 11, D, D, 1, 1), ("relu", 12)]

    in_dim = decoder_channels // r if use_decoder_state_for_postnet_input else mel_dim
    C, Fd = converter_channels, linear_dim
    post = [("conv", 0, in_dim, C, 1, 1), ("hw", 1, C, k, 1, False), ("hw", 2, C, k, 3, False),
            ("convT", 3, C, C), ("hw", 4, C, k, 1, False), ("hw", 5, C, k, 3, False),
            ("convT", 6, C, C), ("hw", 7, C, k, 1, False), ("hw", 8, C, k, 3, False),
            ("conv", 9, C, 2 * C, 1, 1), ("hw", 10, 2 * C, k, 1, False),
            ("hw", 11, 2 * C, k, 1, False), ("conv", 12, 2 * C, Fd, 1, 1),
            ("conv", 13, Fd, Fd, 1, 1), ("relu", 14), ("conv", 15, Fd, Fd, 1, 1), ("relu", 16),
            ("conv", 17, Fd, Fd, 1, 1), ("sigmoid", 18)]
    return dict(kind="nyanko", n_vocab=n_vocab, embed_dim=embed_dim, mel_dim=mel_dim,
                linear_dim=linear_dim, r=r, n_speakers=1, speaker_embed_dim=None,
                padding_idx=padding_idx,
                encoder=enc, audio_encoder=aenc, audio_decoder=adec, converter=post,
                decoder_channels=decoder_channels, use_memory_mask=use_memory_mask,
                use_decoder_state_for_postnet_input=use_decoder_state_for_postnet_input,
                key_projection=key_projection, value_projection=value_projection,
                max_positions=max_positions)


def spec_from_builder(builder_name, **kw):
    if builder_name == "deepvoice3":
        return deepvoice3_spec(multispeaker_recipe=False, **kw)
    if builder_name == "deepvoice3_multispeaker":
        kw.setdefault("key_projection", True)      # builder.py:191-192 defaults
        kw.setdefault("value_projection", True)
        return deepvoice3_spec(multispeaker_recipe=True, **kw)
    if builder_name == "nyanko":
        return nyanko_spec(**kw)
    raise ValueError(builder_name)
