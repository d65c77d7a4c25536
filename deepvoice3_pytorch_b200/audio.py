"""Audio front-end with the reference's function names (reference audio.py): ``spectrogram(y)`` and
``melspectrogram(y)`` take a float waveform and return (n_freq, n_frames) arrays in [0, 1] -- computed by ONE fused
GPU pass (csrc/stft.cu) instead of two CPU lws STFTs.  ``stft_mel_batch`` is the batched device API the
preprocessors should use (a whole shard of clips per launch; one H2D, one D2H).

``inv_spectrogram`` (reference audio.py:37-43) is provided with Griffin-Lim phase recovery on the same STFT frame: the
reference's LWS (``lws`` package) is an un-vendored dependency, so that path's parity is unpinned (csrc/istft.cu).
"""
import ctypes

import numpy as np
import torch

from ._lib import lib, Dv3Error


class _HP:
    """Defaults of reference hparams.py / presets/*.json; override by assigning attributes."""
    sample_rate = 22050
    fft_size = 1024
    hop_size = 256
    num_mels = 80
    fmin = 125
    fmax = 7600
    preemphasis = 0.97
    min_level_db = -100
    ref_level_db = 20
    power = 1.4                   # spectrogram sharpening before phase recovery (presets/*.json)
    griffin_lim_iters = 60
    rescaling = False             # preprocess.py: y = x / |x|.max() * rescaling_max (hparams.py:46-48)
    rescaling_max = 0.999
    min_text = 20                 # utterances with shorter transcripts are skipped (hparams.py:137)


hparams = _HP()
_basis_cache = {}


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f / (200.0 / 3)
    return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) / (np.log(6.4) / 27.0), lin)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), (200.0 / 3) * m)


def _build_mel_basis():
    """Slaney-scale, area-normalised triangular filterbank = librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)
    (reference audio.py:71-76); (num_mels, fft_size//2+1) float32."""
    hp = hparams
    if hp.fmax is not None:
        assert hp.fmax <= hp.sample_rate // 2
    n_bins = 1 + hp.fft_size // 2
    freqs = np.linspace(0, hp.sample_rate / 2.0, n_bins)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(hp.fmin), _hz_to_mel(hp.fmax or hp.sample_rate / 2.0),
                                   hp.num_mels + 2))
    basis = np.zeros((hp.num_mels, n_bins))
    for m in range(hp.num_mels):
        lo, ce, hi = edges[m], edges[m + 1], edges[m + 2]
        up = (freqs - lo) / (ce - lo)
        down = (hi - freqs) / (hi - ce)
        basis[m] = np.maximum(0.0, np.minimum(up, down)) * (2.0 / (hi - lo))
    return basis.astype(np.float32)


def _device_basis(device):
    hp = hparams
    key = (str(device), hp.sample_rate, hp.fft_size, hp.num_mels, hp.fmin, hp.fmax)
    if key not in _basis_cache:
        basis = _build_mel_basis()
        nz = basis > 0
        start = np.array([int(np.argmax(r)) if r.any() else 0 for r in nz], dtype=np.int32)
        length = np.array([int(len(r) - np.argmax(r[::-1]) - s) if r.any() else 0
                           for r, s in zip(nz, start)], dtype=np.int32)
        _basis_cache[key] = (torch.from_numpy(basis).to(device), torch.from_numpy(start).to(device),
                             torch.from_numpy(length).to(device))
    return _basis_cache[key]


def load_wav(path):
    """float32 mono waveform in [-1, 1] at ``hparams.sample_rate`` -- reference audio.py:12-13
    (``librosa.core.load(path, sr=hparams.sample_rate)[0]``).  Host-side file plumbing (scipy): integer PCM is scaled by
    its full range, channels are averaged, and a file at another rate is resampled with a polyphase filter (librosa
    uses resampy's kaiser_best; identical output only when the rates already agree, as for the reference's datasets)."""
    from scipy.io import wavfile
    sr, x = wavfile.read(path)
    if np.issubdtype(x.dtype, np.integer):
        x = x.astype(np.float32) / float(2 ** (8 * x.dtype.itemsize - 1)) if x.dtype != np.uint8 \
            else (x.astype(np.float32) - 128.0) / 128.0
    else:
        x = x.astype(np.float32)
    if x.ndim > 1:
        x = x.mean(axis=1)
    if sr != hparams.sample_rate:
        from math import gcd
        from scipy.signal import resample_poly
        g = gcd(int(sr), int(hparams.sample_rate))
        x = resample_poly(x, hparams.sample_rate // g, sr // g).astype(np.float32)
    return np.ascontiguousarray(x, dtype=np.float32)


def save_wav(wav, path):
    """16-bit PCM at ``hparams.sample_rate``, peak-normalised exactly like reference audio.py:16-18."""
    from scipy.io import wavfile
    wav = np.asarray(wav, dtype=np.float64)
    wav = wav * 32767 / max(0.01, np.max(np.abs(wav)))
    wavfile.write(path, hparams.sample_rate, wav.astype(np.int16))


def preemphasis(x):
    """y[n] = x[n] - c*x[n-1] on the host (reference audio.py:21-23, ``lfilter([1, -c], [1], x)``); the fused kernel
    applies the same filter on the fly, this function exists for callers that want the signal itself."""
    from scipy import signal
    return signal.lfilter([1, -hparams.preemphasis], [1], np.asarray(x))


def _linear_to_mel(spectrogram):
    """mel_basis @ |S| -- reference audio.py:64-68 (host-side helper; the kernel fuses it)."""
    return np.dot(_build_mel_basis(), spectrogram)


def num_frames(n_samples):
    return lib.raw("dv3_stft_num_frames")(int(n_samples))


def stft_mel_batch(wav, lengths=None, want_linear=True, want_mel=True):
    """wav: (nclips, max_len) fp32 CUDA tensor; lengths: int32 CUDA tensor (nclips) or None (= all max_len).
    -> linear (nclips, max_frames, 513), mel (nclips, max_frames, num_mels) in the stored (T, F) layout."""
    if not (torch.is_tensor(wav) and wav.is_cuda and wav.dtype == torch.float32 and wav.dim() == 2):
        raise Dv3Error("stft_mel_batch needs a (nclips, max_len) fp32 CUDA tensor; there is no CPU path")
    if hparams.fft_size != 1024 or hparams.hop_size != 256:
        raise Dv3Error("the fused kernel is built for fft_size=1024, hop_size=256 (every reference preset)")
    wav = wav.contiguous()
    nclips, max_len = wav.shape
    dev = wav.device
    if lengths is None:
        lengths = torch.full((nclips,), max_len, dtype=torch.int32, device=dev)
    lengths = lengths.to(device=dev, dtype=torch.int32).contiguous()
    max_frames = num_frames(max_len)
    basis, start, length = _device_basis(dev)
    lin = torch.empty(nclips, max_frames, hparams.fft_size // 2 + 1, device=dev) if want_linear else None
    mel = torch.empty(nclips, max_frames, hparams.num_mels, device=dev) if want_mel else None   # kernel zero-fills ragged tails

    def p(t):
        return None if t is None else ctypes.c_void_p(t.data_ptr())
    lib.call("dv3_stft_mel", p(wav), p(lengths), p(basis), p(start), p(length), p(lin), p(mel), nclips, max_len,
             max_frames, hparams.num_mels, float(hparams.preemphasis), float(hparams.min_level_db),
             float(hparams.ref_level_db), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    return lin, mel


def _single(y, want_linear, want_mel):
    y = torch.as_tensor(np.asarray(y, dtype=np.float32)).view(1, -1).cuda()
    lin, mel = stft_mel_batch(y, None, want_linear, want_mel)
    return lin, mel


def spectrogram(y):
    """(fft_size//2+1, n_frames) normalised dB magnitude -- reference audio.py:31-34."""
    lin, _ = _single(y, True, False)
    return lin[0].t().cpu().numpy()


def melspectrogram(y):
    """(num_mels, n_frames) normalised dB mel spectrogram -- reference audio.py:46-51."""
    _, mel = _single(y, False, True)
    return mel[0].t().cpu().numpy()


def _amp_to_db(x):
    min_level = np.exp(hparams.min_level_db / 20 * np.log(10))
    return 20 * np.log10(np.maximum(min_level, x))


def _db_to_amp(x):
    return np.power(10.0, x * 0.05)


def _normalize(S):
    return np.clip((S - hparams.min_level_db) / -hparams.min_level_db, 0, 1)


def _denormalize(S):
    return (np.clip(S, 0, 1) * -hparams.min_level_db) + hparams.min_level_db


def _cp(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def inv_num_samples(n_frames):
    """Samples reconstructed from n_frames frames (the hop-aligned length whose forward STFT has n_frames frames)."""
    return (int(n_frames) - 1) * hparams.hop_size - (hparams.fft_size - 2 * hparams.hop_size)


def griffin_lim(mag, n_iter=None):
    """mag: (T, 513) fp32 CUDA tensor of linear magnitudes -> waveform (n,) whose STFT magnitude approximates it.
    x <- istft(mag * exp(i*angle(stft(x)))), started from the zero-phase inverse; every arrow is one kernel launch."""
    if not (torch.is_tensor(mag) and mag.is_cuda and mag.dtype == torch.float32 and mag.dim() == 2):
        raise Dv3Error("griffin_lim needs a (T, 513) fp32 CUDA tensor; there is no CPU path")
    if hparams.fft_size != 1024 or hparams.hop_size != 256 or mag.shape[1] != 513:
        raise Dv3Error("the inverse kernels are built for fft_size=1024, hop_size=256")
    mag = mag.contiguous()
    T = mag.shape[0]
    n = inv_num_samples(T)
    if n < 1:
        raise Dv3Error("too few frames (%d) to reconstruct a waveform" % T)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    spec = torch.zeros(T, 513, 2, device=mag.device)
    spec[..., 0] = mag                                   # zero phase
    x = torch.zeros(n, device=mag.device)
    lib.call("dv3_istft", _cp(spec), _cp(x), n, T, st)
    for _ in range(hparams.griffin_lim_iters if n_iter is None else n_iter):
        lib.call("dv3_stft_complex", _cp(x), n, _cp(mag), _cp(spec), T, st)
        x.zero_()
        lib.call("dv3_istft", _cp(spec), _cp(x), n, T, st)
    return x


def inv_preemphasis(x):
    """y[n] = x[n] + c*y[n-1] -- reference audio.py:26-28.  x: (n,) or (nclips, n) fp32 CUDA tensor -> tensor; a numpy
    array (the reference's calling convention) is moved to the GPU and a numpy array comes back."""
    if not torch.is_tensor(x):
        return inv_preemphasis(torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).cuda()).cpu().numpy()
    x2 = x.view(1, -1) if x.dim() == 1 else x
    x2 = x2.contiguous()
    y = torch.empty_like(x2)
    lib.call("dv3_deemphasis", _cp(x2), _cp(y), x2.shape[0], x2.shape[1], x2.shape[1], float(hparams.preemphasis),
             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    return y.view_as(x)


def inv_spectrogram(spectrogram, n_iter=None):
    """(513, T) normalised dB spectrogram (what ``spectrogram`` returns / the model predicts, transposed) -> waveform
    float32 numpy array -- reference audio.py:37-43: denormalise, dB -> amplitude, ** power, phase recovery, inverse
    STFT, de-emphasis."""
    S = torch.as_tensor(np.ascontiguousarray(np.asarray(spectrogram, dtype=np.float32).T)).cuda()   # (T, 513)
    amp = torch.empty_like(S)
    lib.call("dv3_spec_to_amp", _cp(S), _cp(amp), S.numel(), float(hparams.min_level_db), float(hparams.ref_level_db),
             float(hparams.power), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    return inv_preemphasis(griffin_lim(amp, n_iter)).cpu().numpy()
