"""A/B timing of dv3_stft_mel from two builds of the library in ONE process on one GPU (box-to-box variation is ~10 %,
larger than most kernel changes): tools/ab/libA.so vs tools/ab/libB.so, interleaved, same inputs, outputs compared."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepvoice3_pytorch_b200 import audio

nb, n = 256, 220500
wav = (0.1 * torch.randn(nb, n, device="cuda")).clamp_(-1, 1)
T = audio.num_frames(n)
basis, start, length = audio._device_basis(wav.device)
lens = torch.full((nb,), n, dtype=torch.int32, device="cuda")
p = lambda t: ctypes.c_void_p(t.data_ptr())
outs, libs = {}, {}
for name in sys.argv[1:] or ["A", "B"]:
    dll = ctypes.CDLL(os.path.join(ROOT, "tools", "ab", "lib%s.so" % name))
    f = dll.dv3_stft_mel
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_void_p] * 7 + [ctypes.c_int] * 4 + [ctypes.c_float] * 3 + [ctypes.c_void_p]
    libs[name] = f
    outs[name] = (torch.empty(nb, T, 513, device="cuda"), torch.empty(nb, T, 80, device="cuda"))

def run(name):
    lin, mel = outs[name]
    rc = libs[name](p(wav), p(lens), p(basis), p(start), p(length), p(lin), p(mel), nb, n, T, 80, 0.97, -100.0, 20.0,
                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0

for name in libs:
    for _ in range(3):
        run(name)
torch.cuda.synchronize()
res = {k: [] for k in libs}
for rep in range(5):
    for name in libs:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run(name)
        e1.record(); torch.cuda.synchronize()
        res[name].append(e0.elapsed_time(e1) / 20)
b = 4.0 * nb * (n + T * 513 + T * 80)
for name, v in res.items():
    ms = sorted(v)[len(v) // 2]
    print("%s: median %.4f ms (min %.4f)  %.0f clips/s  %.1f GB/s" % (name, ms, min(v), nb / ms * 1e3, b / ms / 1e6))
names = list(libs)
if len(names) == 2:
    a, c = outs[names[0]], outs[names[1]]
    print("max |dlin| %.3g  max |dmel| %.3g" % ((a[0] - c[0]).abs().max().item(), (a[1] - c[1]).abs().max().item()))
