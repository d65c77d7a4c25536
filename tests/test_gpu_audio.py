"""GPU: fused STFT->linear/mel kernel against the numpy restatement of audio.py (oracle/audio_oracle.py;
parity UNPINNED, see its header).  Outputs are normalised dB in [0,1] (1 unit = 100 dB): tolerance 2e-3 abs
away from the -100 dB clip floor, where fp32-vs-float64 FFT round-off is amplified by the log."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(lin, mel, ref_lin, ref_mel):
    assert lin.shape == ref_lin.shape and mel.shape == ref_mel.shape
    for got, ref in ((lin, ref_lin), (mel, ref_mel)):
        live = ref > 0.05                      # >= 15 dB above the floor
        if live.any():
            assert np.abs(got - ref)[live].max() < 2e-3
        assert np.abs(got - ref).max() < 2e-2
        assert np.abs(got - ref).mean() < 2e-4


def test_single_clip_matches_oracle():
    from deepvoice3_pytorch_b200 import audio
    from oracle import audio_oracle as A
    x = A.synthetic_clip(3)
    ref_lin, ref_mel = A.process_utterance(x)
    lin = audio.spectrogram(x)
    mel = audio.melspectrogram(x)
    assert lin.shape == (513, 865) and mel.shape == (80, 865)
    _check(lin.T, mel.T, ref_lin, ref_mel)


def test_ragged_batch_and_edge_lengths():
    from deepvoice3_pytorch_b200 import audio
    from oracle import audio_oracle as A
    lens = [1, 255, 256, 257, 1023, 1024, 1025, 5000, 22050]
    clips = [A.synthetic_clip(10 + i, n=max(n, 2))[:n] for i, n in enumerate(lens)]
    wav = np.zeros((len(lens), max(lens)), dtype=np.float32)
    for i, c in enumerate(clips):
        wav[i, :len(c)] = c
    lin, mel = audio.stft_mel_batch(torch.from_numpy(wav).cuda(), torch.tensor(lens, dtype=torch.int32).cuda())
    lin, mel = lin.cpu().numpy(), mel.cpu().numpy()
    for i, (c, n) in enumerate(zip(clips, lens)):
        nf = A.num_frames(n)
        assert nf == audio.num_frames(n)
        ref_lin, ref_mel = A.process_utterance(c)
        _check(lin[i, :nf], mel[i, :nf], ref_lin, ref_mel)
        assert not lin[i, nf:].any() and not mel[i, nf:].any()     # untouched beyond the clip


def test_mel_basis_matches_librosa_definition():
    from deepvoice3_pytorch_b200 import audio
    from oracle import audio_oracle as A
    np.testing.assert_allclose(audio._build_mel_basis(), A.mel_basis(), rtol=1e-6, atol=1e-9)


def test_linearity_and_silence():
    """Size-independent properties: silence -> exactly the floor (0); scaling the input by 10 raises every
    unclipped bin by 20 dB (= 0.2 normalised)."""
    from deepvoice3_pytorch_b200 import audio
    from oracle import audio_oracle as A
    z = np.zeros(22050, dtype=np.float32)
    assert not audio.spectrogram(z).any() and not audio.melspectrogram(z).any()
    x = 0.05 * A.synthetic_clip(5, n=44100)
    a, b = audio.spectrogram(x), audio.spectrogram(10 * x)
    ok = (a > 0.05) & (b < 0.95)
    assert np.abs((b - a)[ok] - 0.2).max() < 2e-3
