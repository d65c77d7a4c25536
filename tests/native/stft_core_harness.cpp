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
    // x[-1] must be readable and x[0] 16-byte aligned: four leading floats, the frame starts at xs[4]
    alignas(16) static float xs[1028];
    if (fread(xs + 3, 4, 1025, stdin) != 1025) return 2;
    const float* x = xs + 4;
    std::vector<f4> tab(TAB_N);
    for (int i = 0; i < TAB_N; ++i) tab[i] = table_entry(i);
    const f4 *win = tab.data() + TAB_WIN, *tw1 = tab.data() + TAB_TW1, *tw2 = tab.data() + TAB_TW2,
             *wsp = tab.data() + TAB_WSP;
    const int lim = (int)hdr[1];

    // NaN-poisoned work planes: a stage that reads a word no earlier stage wrote shows up in the output
    alignas(16) static float re[WORK], im[WORK];
    for (int i = 0; i < WORK; ++i) re[i] = im[i] = NAN;
    static pr vr[32][8], vi[32][8];
    for (int l = 0; l < 32; ++l) {
        if (lim < 1024) pass1<true>(l, x, hdr[0], lim, win, tw1, vr[l], vi[l]);
        else pass1<false>(l, x, hdr[0], lim, win, tw1, vr[l], vi[l]);
    }
    for (int l = 0; l < 32; ++l) store1(l, vr[l], vi[l], re, im);
    for (int l = 0; l < 32; ++l) pass2(l, re, im, tw2, vr[l], vi[l]);
    for (int l = 0; l < 32; ++l) store2(l, vr[l], vi[l], re, im);
    for (int l = 0; l < 32; ++l) pass3(l, re, im, vr[l], vi[l]);
    for (int l = 0; l < 32; ++l) store3(l, vr[l], vi[l], re, im);
    std::vector<float> mag(513);
    for (int j = 0; j < 4; ++j)
        for (int l = 0; l < 32; ++l) {
            const int ka = l + 64 * j;
            pr lo, hi;
            split4(ka, re, im, rot16(wsp[l], j), lo, hi);
            mag[ka] = 0.5f * std::sqrt(lo.x);
            mag[ka + 32] = 0.5f * std::sqrt(lo.y);
            mag[512 - ka] = 0.5f * std::sqrt(hi.x);
            mag[512 - ka - 32] = 0.5f * std::sqrt(hi.y);
        }
    mag[256] = 0.5f * std::sqrt(split_nyquist(re, im));
    fwrite(mag.data(), 4, 513, stdout);
    return 0;
}
