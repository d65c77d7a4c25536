"""The drop-in claim, exercised: the reference's OWN ``train.py`` (loaded unchanged from oracle/_ref through
oracle/ref_harness.py, which only stands in for uninstalled CLI / logging / text packages) runs on top of
``deepvoice3_pytorch_b200`` -- ``build_model()``, ``collate_fn`` and the ``train()`` loop with ``model(...)``,
``loss.backward()``, ``clip_grad_norm_`` and ``torch.optim.Adam`` -- and produces the same losses as the same file on
the reference package."""
import os

import numpy as np
import pytest
import torch


def _harness():
    from oracle import ref_harness as H
    if H.ref_root() is None:
        pytest.skip("oracle/_ref not built (python oracle/make_ref.py in the build container)")
    return H


def test_train_py_build_model_on_this_package_cpu():
    """train.py:812-840 ``build_model()`` bound to this package: same class surface, same state_dict keys and -- for
    the same seed -- bit-identical initial weights as on the reference package."""
    H = _harness()
    sds = {}
    for which in ("reference", "b200"):
        tr = H.load_train(which)
        H.apply_preset(tr, "deepvoice3_ljspeech")
        torch.manual_seed(0)
        model = tr.build_model()
        assert hasattr(model, "get_trainable_parameters") and hasattr(model.seq2seq.decoder, "max_decoder_steps")
        sds[which] = model.state_dict()
    assert type(model).__module__.startswith("deepvoice3_pytorch_b200")
    assert list(sds["reference"].keys()) == list(sds["b200"].keys())
    for k in sds["reference"]:
        assert torch.equal(sds["reference"][k], sds["b200"][k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("preset,n_speakers", [("deepvoice3_ljspeech", 1), ("nyanko_ljspeech", 1),
                                               ("deepvoice3_vctk", 108)])
def test_reference_train_loop_runs_unchanged_on_this_package(preset, n_speakers, tmp_path):
    """4 optimizer steps of reference ``train()`` on synthetic utterances batched by reference ``collate_fn``: the
    reference package in PyTorch eager on the same GPU (TF32 off) vs this package (default tensor-core mode, eager --
    what a user gets by pointing ``train.py`` at this package).  dropout = 0 so the two are comparable."""
    H = _harness()
    from torch.utils.data import DataLoader
    logs = {}
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for which in ("reference", "b200"):
            tr = H.load_train(which)
            hp = H.apply_preset(tr, preset, dropout=0.0, eval_interval=10 ** 9)
            torch.manual_seed(0)
            model = tr.build_model().to("cuda")
            opt = torch.optim.Adam(model.get_trainable_parameters(), lr=hp.initial_learning_rate,
                                   betas=(hp.adam_beta1, hp.adam_beta2), eps=hp.adam_eps,
                                   weight_decay=hp.weight_decay, amsgrad=hp.amsgrad)
            utts = H.synthetic_utterances(16, seed=3, n_speakers=n_speakers, min_text=40, max_text=100, min_frames=200,
                                          max_frames=400)
            loader = DataLoader(utts, batch_size=4, collate_fn=tr.collate_fn, shuffle=False)
            writer = H.ScalarLog()
            tr.global_step, tr.global_epoch = 0, 0
            tr.train(torch.device("cuda"), model, loader, opt, writer, init_lr=hp.initial_learning_rate,
                     checkpoint_dir=str(tmp_path), checkpoint_interval=10 ** 9, nepochs=1,
                     clip_thresh=hp.clip_thresh)
            logs[which] = writer.scalars
            del model, opt
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32 = old_tf32
    for tag in ("loss", "mel_l1_loss", "linear_l1_loss", "done_loss", "attn_loss", "gradient norm"):
        ref = np.array([v for _, v in logs["reference"][tag]])
        got = np.array([v for _, v in logs["b200"][tag]])
        assert len(ref) == 4 and len(got) == 4, tag
        np.testing.assert_allclose(got[0], ref[0], rtol=2e-4, err_msg=tag + " (first step: identical weights)")
        np.testing.assert_allclose(got, ref, rtol=5e-3, err_msg=tag)
    assert all(np.isfinite(v) for _, v in logs["b200"]["loss"])


@pytest.mark.gpu
def test_reference_synthesis_tts_runs_unchanged_on_this_package():
    """The inference-side caller: reference ``synthesis.py:tts()`` (text ids -> ``model(...)`` autoregressive decoding
    -> ``audio._denormalize`` / ``audio.inv_spectrogram``) executed UNCHANGED with this package's model and ``audio``
    module, against the same function on the reference package (PyTorch eager on the same GPU; its waveform step uses
    the numpy restatement of audio.py because ``lws`` is not installed).  Same weights, a decoder that never raises the
    done flag early, 24 decoder steps: mel / linear / alignment agree; both waveforms are finite and of the length
    the frame count implies."""
    import types
    H = _harness()
    if not os.path.exists(os.path.join(H.ref_root(), "synthesis.py")):
        pytest.skip("oracle/_ref predates synthesis.py (re-run oracle/make_ref.py)")
    from oracle import audio_oracle as A
    from deepvoice3_pytorch_b200 import audio as our_audio
    shim = types.ModuleType("audio")                       # what the reference side's `import audio` resolves to
    shim._denormalize = lambda S: np.clip(S, 0, 1) * 100.0 - 100.0
    shim.inv_spectrogram = lambda S: A.inv_spectrogram(S, n_iter=4)
    old_iters = our_audio.hparams.griffin_lim_iters
    our_audio.hparams.griffin_lim_iters = 4
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False                # the reference side: exact fp32 convolutions
    outs, sd = {}, None
    try:
        for which, aud in (("reference", shim), ("b200", our_audio)):
            syn = H.load_synthesis(which, aud)
            tr = H.load_train(which)
            H.apply_preset(tr, "deepvoice3_ljspeech")
            torch.manual_seed(0)
            model = tr.build_model()
            if sd is None:
                sd = {k: v.clone() for k, v in model.state_dict().items()}
                sd["seq2seq.decoder.fc.bias"] = torch.full_like(sd["seq2seq.decoder.fc.bias"], -8.0)   # never "done"
            model.load_state_dict(sd)
            model.seq2seq.decoder.max_decoder_steps = 24
            waveform, alignment, spectrogram, mel = syn.tts(model, "hello b200", p=0, speaker_id=None, fast=True)
            outs[which] = (np.asarray(waveform), np.asarray(alignment), np.asarray(spectrogram), np.asarray(mel))
    finally:
        our_audio.hparams.griffin_lim_iters = old_iters
        torch.backends.cudnn.allow_tf32 = old_tf32
    ref, got = outs["reference"], outs["b200"]
    assert type(model).__module__.startswith("deepvoice3_pytorch_b200")
    for name, r, g in zip(("alignment", "spectrogram", "mel"), ref[1:], got[1:]):
        assert r.shape == g.shape, name
        scale = max(1.0, float(np.abs(r).max()))
        np.testing.assert_allclose(g, r, rtol=2e-3, atol=2e-3 * scale, err_msg=name)
    T = ref[2].shape[0]
    assert got[2].shape == (T, 513) and got[3].shape[1] == 80
    n = our_audio.inv_num_samples(T)
    assert got[0].shape == (n,) and np.isfinite(got[0]).all() and np.abs(got[0]).max() > 0
    assert ref[0].shape[0] == n and np.isfinite(ref[0]).all()


@pytest.mark.gpu
def test_reference_ljspeech_preprocessor_and_batched_equivalent(tmp_path):
    """The caller of the audio front-end: reference ``ljspeech.py:_process_utterance`` (load_wav -> spectrogram ->
    melspectrogram -> np.save, :40-79) runs UNCHANGED on this package's ``audio`` module, one clip per launch; the
    batched ``preprocess.build_from_path`` (same arguments, one fused launch per batch) writes bit-identical files and
    returns the same ``train.txt`` rows, including the reference's skip of short transcripts."""
    from scipy.io import wavfile
    H = _harness()
    if not os.path.exists(os.path.join(H.ref_root(), "ljspeech.py")):
        pytest.skip("oracle/_ref predates ljspeech.py (re-run oracle/make_ref.py)")
    from oracle import audio_oracle as A
    from deepvoice3_pytorch_b200 import audio, preprocess
    in_dir, out_a, out_b = tmp_path / "in", tmp_path / "ref", tmp_path / "ours"
    (in_dir / "wavs").mkdir(parents=True)
    out_a.mkdir(); out_b.mkdir()
    lens = [30001, 22050, 5000, 47777, 12345]
    texts = ["the quick brown fox jumps over the lazy dog %d" % i for i in range(5)]
    texts[2] = "too short"                                  # < hparams.min_text = 20: skipped by both
    with open(in_dir / "metadata.csv", "w", encoding="utf-8") as f:
        for i, (n, t) in enumerate(zip(lens, texts)):
            x = A.synthetic_clip(70 + i, n=n)
            wavfile.write(str(in_dir / "wavs" / ("LJ%03d.wav" % i)), 22050, (x * 32767).astype(np.int16))
            f.write("LJ%03d|%s|%s\n" % (i, t, t))
    lj = H.load_ljspeech(audio)
    assert lj.hparams.min_text == audio.hparams.min_text and not lj.hparams.rescaling
    rows_ref, index = [], 1
    for i, t in enumerate(texts):
        if len(t) < lj.hparams.min_text:
            continue
        rows_ref.append(lj._process_utterance(str(out_a), index, str(in_dir / "wavs" / ("LJ%03d.wav" % i)), t))
        index += 1
    rows = preprocess.build_from_path(str(in_dir), str(out_b), num_workers=2, batch_clips=3)
    assert rows == rows_ref and len(rows) == 4
    for spec_name, mel_name, n_frames, _ in rows:
        a, b = np.load(out_a / spec_name), np.load(out_b / spec_name)
        assert a.shape == (n_frames, 513) and a.dtype == np.float32 and np.array_equal(a, b)
        a, b = np.load(out_a / mel_name), np.load(out_b / mel_name)
        assert a.shape == (n_frames, 80) and np.array_equal(a, b)
    frames, hours = preprocess.write_metadata(rows, str(out_b))
    lines = open(out_b / "train.txt", encoding="utf-8").read().splitlines()
    assert len(lines) == 4 and lines[0].split("|")[:3] == ["ljspeech-spec-00001.npy", "ljspeech-mel-00001.npy",
                                                           str(rows[0][2])] and frames == sum(r[2] for r in rows)
    # and against the numpy restatement of audio.py for one utterance
    ref_lin, ref_mel = A.process_utterance(audio.load_wav(str(in_dir / "wavs" / "LJ000.wav")))
    got = np.load(out_b / rows[0][0])
    assert got.shape == ref_lin.shape and np.abs(got - ref_lin).mean() < 2e-4
