// Per-warp 1024-point real STFT frame: the index arithmetic of stft.cu, written so that g++ can compile it too
// (tests/test_stft_core.py runs the stages lane by lane on the CPU and compares with numpy's rfft).
//
// One warp transforms one frame.  1024 real samples are packed as 512 complex points z[n] = x[2n] + i*x[2n+1];
// the 512-point transform is three radix-8 passes (512 = 8*8*8) with the 16 points of a lane held in registers
// (two "virtual threads" t = lane + 32*h of 8 points each) and two exchanges through the warp's shared-memory
// work area -- __syncwarp only, no block barrier.  With n = 64*n2 + 8*n1 + n0 and k = k0 + 8*k1 + 64*k2:
//   pass 1  thread t = 8*n1 + n0 : A[k0]  = W512^(t*k0)  * sum_n2 z[64*n2 + t]      * W8^(n2*k0)
//   pass 2  thread (k0, n0)      : B[k1]  = W64^(n0*k1)  * sum_n1 A[k0][8*n1 + n0]  * W8^(n1*k1)
//   pass 3  thread (k0, k1)      : Z[k]   =                sum_n0 B[k0][k1][n0]     * W8^(n0*k2)
// Exchange layouts are chosen so that every warp-wide access hits 32 distinct banks:
//   exchange 1  addr = 72*k0 + t                 written t-contiguous, read by lanes (k0 = lane/8 + 4h, n0 = lane%8)
//   exchange 2  addr = k0 + 8*k1 + 68*n0         written by those lanes, read by lanes (k0 = lane%8, k1 = lane/8 + 4h)
//   natural     addr = k                         written by those lanes (k mod 32 = k0 + 8*(k1%4)), read contiguously
// The half spectrum follows from Z by the usual even/odd split, two bins (k, 512-k) per step.
#pragma once
#if defined(__CUDACC__)
#define STFT_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define STFT_HD inline
#endif

namespace dv3 {
namespace stftc {

struct cpx { float r, i; };
struct f2 { float x, y; };

constexpr int WORK = 578;        // floats per plane (re / im) of a warp's work area; 2*WORK = 4 (mod 32), see stft.cu
constexpr int PITCH1 = 72;       // exchange 1 row pitch
constexpr int PITCH2 = 68;       // exchange 2 n0 pitch

STFT_HD cpx cadd(cpx a, cpx b) { return {a.r + b.r, a.i + b.i}; }
STFT_HD cpx csub(cpx a, cpx b) { return {a.r - b.r, a.i - b.i}; }
STFT_HD cpx cmul(cpx a, f2 w) { return {a.r * w.x - a.i * w.y, a.r * w.y + a.i * w.x}; }

// in-place 8-point forward DFT: a[k] <- sum_n a[n] * exp(-2*pi*i*n*k/8)
STFT_HD void radix8(cpx* a) {
    const float h = 0.70710678118654752f;
    const cpx b0 = cadd(a[0], a[4]), b1 = csub(a[0], a[4]), b2 = cadd(a[2], a[6]), b3 = csub(a[2], a[6]);
    const cpx b4 = cadd(a[1], a[5]), b5 = csub(a[1], a[5]), b6 = cadd(a[3], a[7]), b7 = csub(a[3], a[7]);
    const cpx e0 = cadd(b0, b2), e2 = csub(b0, b2);
    const cpx e1 = {b1.r + b3.i, b1.i - b3.r}, e3 = {b1.r - b3.i, b1.i + b3.r};          // b1 -/+ i*b3
    const cpx o0 = cadd(b4, b6), o2 = csub(b4, b6);
    const cpx o1 = {b5.r + b7.i, b5.i - b7.r}, o3 = {b5.r - b7.i, b5.i + b7.r};
    const cpx t1 = {h * (o1.r + o1.i), h * (o1.i - o1.r)};                               // (1-i)/sqrt2 * o1
    const cpx t2 = {o2.i, -o2.r};                                                        // -i * o2
    const cpx t3 = {h * (o3.i - o3.r), -h * (o3.r + o3.i)};                              // (-1-i)/sqrt2 * o3
    a[0] = cadd(e0, o0); a[4] = csub(e0, o0);
    a[1] = cadd(e1, t1); a[5] = csub(e1, t1);
    a[2] = cadd(e2, t2); a[6] = csub(e2, t2);
    a[3] = cadd(e3, t3); a[7] = csub(e3, t3);
}

// pass 1: x -> sample 0 of the frame window in the RAW waveform (x[-1] is readable; samples outside the clip are 0);
// pre-emphasis e[i] = x[i] - c*x[i-1] (audio.py:21-23) is applied on the fly and e[i] = 0 for i >= lim (the padding
// after the clip's last sample); win[n] = (w[2n], w[2n+1]), tw1[(k0-1)*64 + t] = W512^(t*k0)
STFT_HD void pass1(int lane, const float* x, float preemph, int lim, const f2* win, const f2* tw1, cpx (&v)[2][8]) {
    const f2* xz = reinterpret_cast<const f2*>(x);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int t = lane + 32 * h;
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) {
            const int n = 64 * n2 + t;
            const f2 s = xz[n], w = win[n];
            const float xm = x[2 * n - 1];
            float e0 = s.x - preemph * xm, e1 = s.y - preemph * s.x;
            if (lim < 1024) {                                    // warp-uniform: only a clip's last frames
                if (2 * n >= lim) e0 = 0.f;
                if (2 * n + 1 >= lim) e1 = 0.f;
            }
            v[h][n2] = {e0 * w.x, e1 * w.y};
        }
        radix8(v[h]);
#pragma unroll
        for (int k0 = 1; k0 < 8; ++k0) v[h][k0] = cmul(v[h][k0], tw1[(k0 - 1) * 64 + t]);
    }
}
STFT_HD void store1(int lane, const cpx (&v)[2][8], float* re, float* im) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k0 = 0; k0 < 8; ++k0) {
            re[PITCH1 * k0 + lane + 32 * h] = v[h][k0].r;
            im[PITCH1 * k0 + lane + 32 * h] = v[h][k0].i;
        }
}
// pass 2: tw2[(k1-1)*8 + n0] = W64^(n0*k1)
STFT_HD void pass2(int lane, const float* re, const float* im, const f2* tw2, cpx (&v)[2][8]) {
    const int n0 = lane & 7;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k0 = (lane >> 3) + 4 * h;
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) v[h][n1] = {re[PITCH1 * k0 + 8 * n1 + n0], im[PITCH1 * k0 + 8 * n1 + n0]};
        radix8(v[h]);
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) v[h][k1] = cmul(v[h][k1], tw2[(k1 - 1) * 8 + n0]);
    }
}
STFT_HD void store2(int lane, const cpx (&v)[2][8], float* re, float* im) {
    const int n0 = lane & 7;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k0 = (lane >> 3) + 4 * h;
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) {
            re[k0 + 8 * k1 + PITCH2 * n0] = v[h][k1].r;
            im[k0 + 8 * k1 + PITCH2 * n0] = v[h][k1].i;
        }
    }
}
STFT_HD void pass3(int lane, const float* re, const float* im, cpx (&v)[2][8]) {
    const int k0 = lane & 7;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k1 = (lane >> 3) + 4 * h;
#pragma unroll
        for (int n0 = 0; n0 < 8; ++n0) v[h][n0] = {re[k0 + 8 * k1 + PITCH2 * n0], im[k0 + 8 * k1 + PITCH2 * n0]};
        radix8(v[h]);
    }
}
STFT_HD void store3(int lane, const cpx (&v)[2][8], float* re, float* im) {      // natural order Z[k], k < 512
    const int k0 = lane & 7;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k1 = (lane >> 3) + 4 * h;
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
            re[k0 + 8 * k1 + 64 * k2] = v[h][k2].r;
            im[k0 + 8 * k1 + 64 * k2] = v[h][k2].i;
        }
    }
}
// bins k and 512-k (0 <= k <= 256) from Z[k], Z[512-k]; w = W1024^k.  Returns |2*X[k]|^2 and |2*X[512-k]|^2:
//   2E = Z[k] + conj(Z[512-k]),  2O = -i*(Z[k] - conj(Z[512-k])),  X[k] = E + w*O,  X[512-k] = conj(E - w*O)
STFT_HD void split_pair(int k, const float* re, const float* im, f2 w, float& p_lo, float& p_hi) {
    const int ka = k & 511, kb = (512 - k) & 511;
    const float ar = re[ka], ai = im[ka], br = re[kb], bi = -im[kb];
    const float er = ar + br, ei = ai + bi, dr = ar - br, di = ai - bi;
    const float orr = di, oi = -dr;
    const float pr = w.x * orr - w.y * oi, pi = w.x * oi + w.y * orr;
    const float xr = er + pr, xi = ei + pi, yr = er - pr, yi = ei - pi;
    p_lo = xr * xr + xi * xi;
    p_hi = yr * yr + yi * yi;
}

}  // namespace stftc
}  // namespace dv3
