"""CPU: the index arithmetic of the STFT kernel (csrc/stft_core.cuh: three radix-8 passes, two warp exchanges, the
even/odd split) compiled with g++ and run lane by lane, against numpy's rfft of the same windowed frame."""
import os
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("stft") / "stft_core_harness")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "deepvoice3_pytorch_b200", "csrc"),
                           os.path.join(ROOT, "tests", "native", "stft_core_harness.cpp"), "-o", exe])
    return exe


@pytest.mark.parametrize("seed,lim", [(0, 1024), (1, 1024), (2, 1024), (3, 700), (4, 1)])
def test_frame_magnitudes_match_rfft(harness, seed, lim):
    """x[-1..1023] raw samples; the harness applies pre-emphasis 0.97 and zeroes the window from sample ``lim`` on
    (the padding after a clip's end), exactly as the kernel's first pass does."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(1025).astype(np.float32)
    if seed == 2:
        x[:] = 0; x[4] = 1.0                      # impulse: flat spectrum, exercises every twiddle
    c = np.float32(0.97)
    hdr = np.array([c, lim], dtype=np.float32)
    out = subprocess.run([harness], input=hdr.tobytes() + x.tobytes(), stdout=subprocess.PIPE, check=True).stdout
    got = np.frombuffer(out, dtype=np.float32)
    e = x[1:].astype(np.float64) - float(c) * x[:-1].astype(np.float64)
    e[lim:] = 0
    i = np.arange(1024)
    w = np.sqrt((0.5 - 0.5 * np.cos(2 * np.pi * (i + 0.5) / 1024)) * 0.5)
    ref = np.abs(np.fft.rfft(e * w))
    assert got.shape == (513,)
    np.testing.assert_allclose(got, ref, rtol=2e-4, atol=2e-5 * ref.max())
