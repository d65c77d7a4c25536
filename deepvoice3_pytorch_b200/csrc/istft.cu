// Inverse audio path: reference audio.py:37-43 (inv_spectrogram) = _denormalize (:92-93) -> _db_to_amp (:84-85) ->
// ** power -> phase recovery -> inverse STFT -> inv_preemphasis (:26-28).
// The reference recovers the phase with the `lws` package (Local Weighted Sums), an un-vendored, unpinned dependency
// whose source is absent -- PARITY UNPINNED.  What is restated here is the published Griffin-Lim fixed-point iteration
// on the SAME STFT frame the forward path uses (sqrt-Hann window * sqrt(2*hop/fsize), 1024 / hop 256, 768 samples of
// zero padding on both sides = lws "perfectrec": sum of squared windows over the 4 overlapping frames == 1, so the
// synthesis window equals the analysis window and no normalisation pass is needed):
//     x <- istft(S * exp(i * angle(stft(x))))        (host loop in audio.py; each arrow below is one launch)
// Kernels (one CTA of 256 threads per frame, the shared-memory radix-2 transform of stft.cu):
//   spec_to_amp_kernel      normalised dB spectrogram -> linear magnitude ** power
//   stft_complex_kernel     waveform -> complex half spectrum (frames, 513) [optionally projected onto a magnitude]
//   istft_kernel            complex half spectrum -> windowed frame, overlap-added into the waveform (atomicAdd)
//   deemphasis_kernel       y[n] = x[n] + c*y[n-1] (a 1st-order IIR: one thread per clip, chunks staged through smem)
#include "common.cuh"

namespace dv3 {

constexpr int IFFT_N = 1024, IHOP = 256, INH = 512, INBINS = 513, IPAD = IFFT_N - IHOP;

__device__ __forceinline__ int ibitrev9(int x) { return (int)(__brev((unsigned)x) >> 23); }
__device__ __forceinline__ float frame_window(int i) {           // sqrt(hann(i) * 2*hop/N), hann = .5*(1-cos(2pi(i+.5)/N))
    const float hann = 0.5f - 0.5f * cospif((2 * i + 1) / (float)IFFT_N);
    return sqrtf(hann * (2.f * IHOP / IFFT_N));
}

// 512-point complex radix-2 DIT FFT in shared memory (input in bit-reversed order), 256 threads, forward sign
__device__ __forceinline__ void fft512(float* zr, float* zi, const float* twr, const float* twi, int tid) {
#pragma unroll
    for (int s = 0; s < 9; ++s) {
        const int half = 1 << s;
        const int pos = tid & (half - 1);
        const int i0 = ((tid >> s) << (s + 1)) + pos, i1 = i0 + half;
        const int tw = pos << (8 - s);
        const float wr = twr[tw], wi = twi[tw];
        const float ar = zr[i0], ai = zi[i0], br0 = zr[i1], bi0 = zi[i1];
        const float br = br0 * wr - bi0 * wi, bi = br0 * wi + bi0 * wr;
        zr[i0] = ar + br; zi[i0] = ai + bi;
        zr[i1] = ar - br; zi[i1] = ai - bi;
        __syncthreads();
    }
}

// S (n) in [0,1] (normalised dB, audio.py:88-89) -> amplitude ** power:  dB = S*(-min_db) + min_db + ref_db
__global__ void spec_to_amp_kernel(const float* __restrict__ s, float* __restrict__ amp, long long n, float min_db,
                                   float ref_db, float power) {
    pdl_trigger(); pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = fminf(fmaxf(s[i], 0.f), 1.f);
        const float db = v * -min_db + min_db + ref_db;                    // audio.py:92-93, :39
        amp[i] = powf(powf(10.f, db * 0.05f), power);                     // audio.py:84-85, :41
    }
}

// wav (len) -> spec (nframes, 513, 2).  mag != null: the result is projected onto that magnitude (Griffin-Lim step):
// spec = mag * X / |X| (X == 0 keeps phase 0).  No preemphasis here (the iteration runs on the pre-emphasised signal).
__global__ void __launch_bounds__(256) stft_complex_kernel(const float* __restrict__ x, int len,
                                                           const float* __restrict__ mag, float* __restrict__ spec,
                                                           int nframes) {
    pdl_trigger(); pdl_wait();
    __shared__ float zr[INH], zi[INH], twr[INH / 2], twi[INH / 2];
    const int frame = blockIdx.x, tid = threadIdx.x;
    if (frame >= nframes) return;
    { float s, c; sincospif(-(float)tid / 256.f, &s, &c); twr[tid] = c; twi[tid] = s; }
    const int base = frame * IHOP - IPAD;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int n = tid + h * 256;
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = 2 * n + e, sidx = base + i;
            v[e] = (sidx >= 0 && sidx < len) ? x[sidx] * frame_window(i) : 0.f;
        }
        const int r = ibitrev9(n);
        zr[r] = v[0]; zi[r] = v[1];
    }
    __syncthreads();
    fft512(zr, zi, twr, twi, tid);
    for (int k = tid; k <= INH; k += 256) {
        const int ka = k & (INH - 1), kb = (INH - k) & (INH - 1);
        const float ar = zr[ka], ai = zi[ka], br = zr[kb], bi = -zi[kb];
        const float er = 0.5f * (ar + br), ei = 0.5f * (ai + bi), dr = 0.5f * (ar - br), di = 0.5f * (ai - bi);
        const float orr = di, oi = -dr;
        float s, c;
        sincospif(-(float)k / 512.f, &s, &c);
        float xr = er + c * orr - s * oi, xi = ei + c * oi + s * orr;
        const size_t o = ((size_t)frame * INBINS + k) * 2;
        if (mag) {
            const float m = mag[(size_t)frame * INBINS + k], a = sqrtf(xr * xr + xi * xi);
            if (a > 0.f) { xr *= m / a; xi *= m / a; } else { xr = m; xi = 0.f; }
        }
        spec[o] = xr; spec[o + 1] = xi;
    }
}

// spec (nframes, 513, 2) -> y (len) += window * irfft(spec[frame]) placed at frame*hop - pad   (y zeroed by the caller)
// Inverse real FFT through the same 512-point complex transform: Z[k] = E[k] + i*O[k] with
// E = (X[k] + conj(X[512-k]))/2, O = (X[k] - conj(X[512-k]))/2 * conj(W1024^k); z = IFFT512(Z); x[2n] = Re z, x[2n+1] = Im z.
// IFFT via conjugation: ifft(Z) = conj(fft(conj(Z))) / 512.
__global__ void __launch_bounds__(256) istft_kernel(const float* __restrict__ spec, float* __restrict__ y, int len,
                                                    int nframes) {
    pdl_trigger(); pdl_wait();
    __shared__ float zr[INH], zi[INH], twr[INH / 2], twi[INH / 2];
    const int frame = blockIdx.x, tid = threadIdx.x;
    if (frame >= nframes) return;
    { float s, c; sincospif(-(float)tid / 256.f, &s, &c); twr[tid] = c; twi[tid] = s; }
    const float* X = spec + (size_t)frame * INBINS * 2;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k = tid + h * 256;                                   // 0..511
        const float ar = X[2 * k], ai = X[2 * k + 1];
        const float br = X[2 * (INH - k)], bi = -X[2 * (INH - k) + 1];    // conj(X[512-k])
        const float er = 0.5f * (ar + br), ei = 0.5f * (ai + bi), dr = 0.5f * (ar - br), di = 0.5f * (ai - bi);
        float s, c;
        sincospif((float)k / 512.f, &s, &c);                           // conj(W1024^k) = exp(+2 pi i k / 1024)
        const float orr = dr * c - di * s, oi = dr * s + di * c;
        // Z = E + i*O ; feed conj(Z) to the forward transform
        const float Zr = er - oi, Zi = ei + orr;
        const int r = ibitrev9(k);
        zr[r] = Zr; zi[r] = -Zi;
    }
    __syncthreads();
    fft512(zr, zi, twr, twi, tid);
    const int base = frame * IHOP - IPAD;
    const float inv = 1.f / 512.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int n = tid + h * 256;
        const float v0 = zr[n] * inv, v1 = -zi[n] * inv;                 // conj back
        const int s0 = base + 2 * n, s1 = s0 + 1;
        if (s0 >= 0 && s0 < len) atomicAdd(&y[s0], v0 * frame_window(2 * n));
        if (s1 >= 0 && s1 < len) atomicAdd(&y[s1], v1 * frame_window(2 * n + 1));
    }
}

// y[n] = x[n] + c*y[n-1]  (audio.py:26-28: lfilter([1], [1, -c], x)).  One CTA per clip: the recurrence is serial, so
// thread 0 walks it while the whole block streams 1024-sample chunks through shared memory (coalesced HBM traffic).
__global__ void __launch_bounds__(256) deemphasis_kernel(const float* __restrict__ x, float* __restrict__ y, int len,
                                                         long long stride, float c) {
    pdl_trigger(); pdl_wait();
    __shared__ float buf[1024];
    const float* xi = x + blockIdx.x * stride;
    float* yo = y + blockIdx.x * stride;
    float prev = 0.f;
    for (int c0 = 0; c0 < len; c0 += 1024) {
        const int n = min(1024, len - c0);
        for (int i = threadIdx.x; i < n; i += 256) buf[i] = xi[c0 + i];
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 0; i < n; ++i) { prev = fmaf(c, prev, buf[i]); buf[i] = prev; }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += 256) yo[c0 + i] = buf[i];
        __syncthreads();
    }
}

}  // namespace dv3

using namespace dv3;

extern "C" {

int dv3_spec_to_amp(const float* spec_norm, float* amp, long long n, float min_level_db, float ref_level_db,
                    float power, void* stream) {
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    launch_k(spec_to_amp_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, spec_norm, amp, n, min_level_db, ref_level_db,
             power);
    return check_launch("spec_to_amp");
}

int dv3_stft_complex(const float* wav, int n_samples, const float* mag, float* spec, int nframes, void* stream) {
    DV3_REQUIRE(nframes >= 1 && n_samples >= 1, "stft_complex: empty input");
    launch_k(stft_complex_kernel, nframes, 256, 0, (cudaStream_t)stream, wav, n_samples, mag, spec, nframes);
    return check_launch("stft_complex");
}

int dv3_istft(const float* spec, float* wav, int n_samples, int nframes, void* stream) {
    DV3_REQUIRE(nframes >= 1 && n_samples >= 1, "istft: empty input");
    launch_k(istft_kernel, nframes, 256, 0, (cudaStream_t)stream, spec, wav, n_samples, nframes);
    return check_launch("istft");
}

int dv3_deemphasis(const float* x, float* y, int nclips, int n_samples, long long stride, float coef, void* stream) {
    DV3_REQUIRE(nclips >= 1 && n_samples >= 1, "deemphasis: empty input");
    launch_k(deemphasis_kernel, nclips, 256, 0, (cudaStream_t)stream, x, y, n_samples, stride, coef);
    return check_launch("deemphasis");
}

}  // extern "C"
