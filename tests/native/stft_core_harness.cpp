// CPU harness for csrc/stft_core.cuh: runs the per-warp stages of one STFT frame lane by lane (each stage for all 32
// lanes before the next = the __syncwarp points of the kernel) and prints |X[k]|, k = 0..512, as raw floats.
// stdin: float32 preemph, float32 lim, then 1025 float32 raw samples x[-1..1023].  Built and driven by
// tests/test_stft_core.py.
#include <cstdio>
#include <vector>
#include "stft_core.cuh"
using namespace dv3::stftc;

int main() {
    float hdr[2];
    if (fread(hdr, 4, 2, stdin) != 2) return 2;
    // x[-1] must be readable and x[0] 8-byte aligned: two leading floats, the frame starts at xs[2]
    static float xs[1026] __attribute__((aligned(16)));
    if (fread(xs + 1, 4, 1025, stdin) != 1025) return 2;
    const float* x = xs + 2;
    const double PI = 3.14159265358979323846;
    std::vector<f2> win(512), tw1(7 * 64), tw2(7 * 8), wsp(257);
    auto w = [&](int i) { return (float)std::sqrt((0.5 - 0.5 * std::cos(2 * PI * (i + 0.5) / 1024)) * 0.5); };
    for (int n = 0; n < 512; ++n) win[n] = {w(2 * n), w(2 * n + 1)};
    for (int k0 = 1; k0 < 8; ++k0)
        for (int t = 0; t < 64; ++t)
            tw1[(k0 - 1) * 64 + t] = {(float)std::cos(-2 * PI * t * k0 / 512), (float)std::sin(-2 * PI * t * k0 / 512)};
    for (int k1 = 1; k1 < 8; ++k1)
        for (int n0 = 0; n0 < 8; ++n0)
            tw2[(k1 - 1) * 8 + n0] = {(float)std::cos(-2 * PI * n0 * k1 / 64), (float)std::sin(-2 * PI * n0 * k1 / 64)};
    for (int k = 0; k <= 256; ++k) wsp[k] = {(float)std::cos(-2 * PI * k / 1024), (float)std::sin(-2 * PI * k / 1024)};

    std::vector<float> re(WORK, 0.f), im(WORK, 0.f);
    static cpx v[32][2][8];
    for (int l = 0; l < 32; ++l) pass1(l, x, hdr[0], (int)hdr[1], win.data(), tw1.data(), v[l]);
    for (int l = 0; l < 32; ++l) store1(l, v[l], re.data(), im.data());
    for (int l = 0; l < 32; ++l) pass2(l, re.data(), im.data(), tw2.data(), v[l]);
    for (int l = 0; l < 32; ++l) store2(l, v[l], re.data(), im.data());
    for (int l = 0; l < 32; ++l) pass3(l, re.data(), im.data(), v[l]);
    for (int l = 0; l < 32; ++l) store3(l, v[l], re.data(), im.data());
    std::vector<float> mag(513);
    for (int k = 0; k <= 256; ++k) {
        float lo, hi;
        split_pair(k, re.data(), im.data(), wsp[k], lo, hi);
        mag[k] = 0.5f * std::sqrt(lo);
        mag[512 - k] = 0.5f * std::sqrt(hi);
        if (k == 256) mag[256] = 0.5f * std::sqrt(lo);
    }
    fwrite(mag.data(), 4, 513, stdout);
    return 0;
}
