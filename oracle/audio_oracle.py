"""TEST INFRASTRUCTURE ONLY.  numpy restatement of the reference's STFT -> linear / mel front-end
(reference audio.py:21-23 preemphasis, :31-34 spectrogram, :46-51 melspectrogram, :54-55 _lws_processor,
:64-93 mel basis, dB, normalisation) at the presets' hparams (fft 1024, hop 256, 22.05 kHz, 80 mels 125-7600 Hz,
preemphasis 0.97, min_level_db -100, ref_level_db 20).

PARITY UNPINNED.  The arithmetic of this path lives in third-party packages that are neither vendored under
/root/reference nor installed here (no network): ``lws`` (unpinned, reference setup.py:87; call sites
audio.py:32,47,55), ``librosa`` (unpinned, setup.py:85; audio.py:74-76) and ``nnmnkwii>=0.0.19`` (setup.py:95;
audio.py:22-28).  The reference's tests pin no STFT/mel values (tests/test_audio.py only round-trips dB<->amp
and is local_only).  What is restated below is those libraries' PUBLISHED algorithms as recalled:

* nnmnkwii.preprocessing.preemphasis(x, c) = scipy.signal.lfilter([1, -c], [1], x): y[n] = x[n] - c*x[n-1], y[0]=x[0].
* lws.lws(fsize, fshift, mode="speech").stft(x): analysis window sqrt(hann(fsize, symmetric) * 2*fshift/fsize)
  with lws's symmetric Hann w[n] = 0.5*(1 - cos(2*pi*(n+0.5)/fsize)) [ASSUMPTION: half-sample offset form];
  "perfectrec" zero padding of fsize-fshift samples on both sides; n_frames = ceil((len_padded - fsize)/fshift)+1
  with the tail zero-padded; rfft of each windowed frame in float64 -> (n_frames, fsize/2+1) complex128.
  Corroboration available here: LJ001-0001 has 212 893 samples and the reference's mel fixture
  (tests/data/ljspeech-mel-00001.npy) has 835 frames = ceil((212893 + 2*768 - 1024)/256) + 1.
* librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with the defaults htk=False, norm='slaney': Slaney mel scale
  (linear below 1 kHz, log above), triangular filters, area normalisation 2/(f[m+2]-f[m]); float32 result.
  Checked in tests against torchaudio.functional.melscale_fbanks(norm="slaney", mel_scale="slaney").
"""
import numpy as np

HP = dict(sample_rate=22050, fft_size=1024, hop_size=256, num_mels=80, fmin=125, fmax=7600,
          preemphasis=0.97, min_level_db=-100, ref_level_db=20)


def preemphasis(x, coef=0.97):
    """reference audio.py:21-23."""
    x = np.asarray(x, dtype=np.float64)
    y = x.copy()
    y[1:] -= coef * x[:-1]
    return y


def lws_window(fsize=1024, fshift=256):
    n = np.arange(fsize, dtype=np.float64)
    hann = 0.5 * (1.0 - np.cos(2.0 * np.pi * (n + 0.5) / fsize))
    return np.sqrt(hann * 2.0 * fshift / fsize)


def num_frames(n_samples, fsize=1024, fshift=256):
    padded = n_samples + 2 * (fsize - fshift)
    return int(np.ceil((padded - fsize) / float(fshift))) + 1


def lws_stft(x, fsize=1024, fshift=256):
    """-> (n_frames, fsize/2+1) complex128."""
    pad = fsize - fshift
    x = np.concatenate([np.zeros(pad), np.asarray(x, dtype=np.float64), np.zeros(pad)])
    M = int(np.ceil((len(x) - fsize) / float(fshift))) + 1
    need = (M - 1) * fshift + fsize
    x = np.concatenate([x, np.zeros(need - len(x))])
    idx = np.arange(fsize)[None, :] + fshift * np.arange(M)[:, None]
    return np.fft.rfft(x[idx] * lws_window(fsize, fshift)[None, :], axis=1)


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_basis(sr=22050, n_fft=1024, n_mels=80, fmin=125, fmax=7600):
    """reference audio.py:71-76 -> librosa.filters.mel: (n_mels, n_fft/2+1) float32."""
    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0, np.minimum(lower, upper))
    weights *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return weights.astype(np.float32)


def _amp_to_db(x, min_level_db=-100):
    """reference audio.py:79-81."""
    min_level = np.exp(min_level_db / 20.0 * np.log(10))
    return 20 * np.log10(np.maximum(min_level, x))


def _normalize(S, min_level_db=-100):
    """reference audio.py:88-89."""
    return np.clip((S - min_level_db) / -min_level_db, 0, 1)


def spectrogram(y, hp=HP):
    """reference audio.py:31-34 -> (fft/2+1, n_frames) float64 in [0,1]."""
    D = lws_stft(preemphasis(y, hp["preemphasis"]), hp["fft_size"], hp["hop_size"]).T
    return _normalize(_amp_to_db(np.abs(D), hp["min_level_db"]) - hp["ref_level_db"], hp["min_level_db"])


def melspectrogram(y, hp=HP):
    """reference audio.py:46-51 -> (num_mels, n_frames) float64 in [0,1]."""
    D = lws_stft(preemphasis(y, hp["preemphasis"]), hp["fft_size"], hp["hop_size"]).T
    basis = mel_basis(hp["sample_rate"], hp["fft_size"], hp["num_mels"], hp["fmin"], hp["fmax"])
    S = _amp_to_db(np.dot(basis, np.abs(D)), hp["min_level_db"]) - hp["ref_level_db"]
    return _normalize(S, hp["min_level_db"])


def process_utterance(wav, hp=HP):
    """What ljspeech.py:63-73 stores per clip: (T, 513) and (T, 80) float32."""
    return spectrogram(wav, hp).astype(np.float32).T, melspectrogram(wav, hp).astype(np.float32).T


def synthetic_clip(seed, n=220500, sr=22050):
    """Seeded 10 s test clip: a few chirping partials + noise, in [-1, 1] (SURVEY.md section 8d metric 3)."""
    rng = np.random.RandomState(seed)
    t = np.arange(n) / float(sr)
    x = 0.1 * rng.randn(n)
    for _ in range(4):
        f0, f1, a = rng.uniform(80, 4000), rng.uniform(80, 4000), rng.uniform(0.05, 0.3)
        x += a * np.sin(2 * np.pi * (f0 * t + 0.5 * (f1 - f0) * t * t / t[-1]))
    env = 0.5 * (1 + np.sin(2 * np.pi * 0.7 * t + rng.uniform(0, 6.28)))
    return np.clip(x * env, -1, 1).astype(np.float32)


# ---- inverse path (reference audio.py:37-43, :26-28).  PARITY UNPINNED: the reference's phase recovery is lws.run_lws
# (source absent); this is the Griffin-Lim iteration the CUDA path implements, restated with numpy rfft / irfft.
def lws_istft(spec, fsize=1024, fshift=256):
    """(n_frames, fsize/2+1) complex -> waveform of (n_frames-1)*fshift - (fsize - 2*fshift) samples (synthesis window
    = analysis window: their squares sum to 1 over the overlapping frames)."""
    M = spec.shape[0]
    pad = fsize - fshift
    frames = np.fft.irfft(spec, n=fsize, axis=1) * lws_window(fsize, fshift)[None, :]
    y = np.zeros((M - 1) * fshift + fsize)
    for m in range(M):
        y[m * fshift:m * fshift + fsize] += frames[m]
    n = (M - 1) * fshift - (fsize - 2 * fshift)
    return y[pad:pad + n]


def griffin_lim(mag, n_iter=60):
    mag = np.asarray(mag, dtype=np.float64)
    x = lws_istft(mag.astype(np.complex128))
    for _ in range(n_iter):
        X = lws_stft(x)[:mag.shape[0]]
        a = np.abs(X)
        X = np.where(a > 0, mag * X / np.maximum(a, 1e-300), mag)
        x = lws_istft(X)
    return x


def inv_preemphasis(x, coef=0.97):
    """scipy.signal.lfilter([1], [1, -coef], x): y[n] = x[n] + coef*y[n-1]."""
    y = np.empty(len(x), dtype=np.float64)
    prev = 0.0
    for i, v in enumerate(np.asarray(x, dtype=np.float64)):
        prev = v + coef * prev
        y[i] = prev
    return y


def inv_spectrogram(spectrogram, power=1.4, n_iter=60):
    """(513, T) normalised dB -> waveform (reference audio.py:37-43 with Griffin-Lim in place of lws.run_lws)."""
    S = np.clip(np.asarray(spectrogram, dtype=np.float64), 0, 1) * -HP["min_level_db"] + HP["min_level_db"]
    amp = np.power(10.0, (S + HP["ref_level_db"]) * 0.05)
    return inv_preemphasis(griffin_lim(amp.T ** power, n_iter), HP["preemphasis"])
